// PostCal above the GPU engine: the reference's class boundary (postcal.h:58-338) with the hot path
// (computeTotalLikelihood, the scoring half of sss_computeTotalLikelihood) delegated to the C-ABI.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <numeric>
#include <random>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "pipsort_host.h"

namespace pipsort_host {

namespace {

// postcal.h:102-112
double addlogSpace(double a, double b) {
    if (a == 0) return b;
    if (b == 0) return a;
    const double base = std::max(a, b);
    if (base - std::min(a, b) > 700) return base;
    return base + std::log(1 + std::exp(std::min(a, b) - base));
}

// postcal.h:277-283
double special_exp(double post, double total) { return post == 0 ? 0.0 : std::exp(post - total); }

void check(int rc) {
    if (rc == 0) return;
    // the reference prints and exits (pipsort.cpp:33-36, postcal.cpp:20-23,291-294); keep its exit codes
    std::cout << pipsort_last_error() << std::endl;
    std::exit(rc == PIPSORT_E_SINGULAR ? 0 : 1);
}

struct ItemByProb {   // util.h data / by_number: descending probability
    double p;
    int idx;
};

}  // namespace

PostCal::PostCal(const std::vector<std::vector<double>>& sigma_eff, const std::vector<std::vector<double>>& z, double K,
                 bool do_sss_, int MAX_causal, const std::vector<std::vector<std::string>>* names, double sharing_param,
                 double gamma, double t_squared, double s_squared, const std::vector<int>& sample_sizes,
                 const std::vector<int>& num_snps, const std::vector<std::vector<int>>& idx_to_snp_map,
                 const std::vector<std::string>& snp_pos, int device, const std::string& configsFile_, int num_configs_,
                 int num_groups_)
    : num_of_studies((int)num_snps.size()), maxCausalSNP(MAX_causal), do_sss(do_sss_), num_snps_all(num_snps),
      SNP_NAME(names), all_snp_pos(snp_pos), configsFile(configsFile_), num_configs(num_configs_), num_groups(num_groups_) {
    totalSnpCount = std::accumulate(num_snps_all.begin(), num_snps_all.end(), 0);
    unionSnpCount = (int)all_snp_pos.size();
    postValues.assign(totalSnpCount, 0.0);
    noCausal.assign(num_of_studies, 0.0);
    sharedPips.assign(unionSnpCount, 0.0);
    sharedLL.assign(unionSnpCount, 0.0);
    notSharedLL.assign(unionSnpCount, 0.0);
    // postcal.cpp:66,89:  d_s = s^2 * double(n_s) / int(min n) + t^2
    const int min_size = *std::min_element(sample_sizes.begin(), sample_sizes.end());
    std::vector<double> d(num_of_studies);
    for (int s = 0; s < num_of_studies; s++) d[s] = s_squared * ((double)sample_sizes[s] / min_size) + t_squared;
    std::vector<double> sig_cat, z_cat;
    for (int s = 0; s < num_of_studies; s++) {
        sig_cat.insert(sig_cat.end(), sigma_eff[s].begin(), sigma_eff[s].end());
        z_cat.insert(z_cat.end(), z[s].begin(), z[s].end());
    }
    std::vector<int32_t> smap((size_t)num_of_studies * unionSnpCount);
    for (int s = 0; s < num_of_studies; s++)
        for (int g = 0; g < unionSnpCount; g++) smap[(size_t)s * unionSnpCount + g] = idx_to_snp_map[s][g];
    std::vector<int32_t> ns(num_snps_all.begin(), num_snps_all.end());
    pipsort_locus loc;
    loc.num_studies = num_of_studies;
    loc.num_snps = ns.data();
    loc.sigma = sig_cat.data();
    loc.z = z_cat.data();
    loc.d = d.data();
    loc.K = K;
    loc.union_count = unionSnpCount;
    loc.snp_map = smap.data();
    loc.gamma = gamma;
    loc.sharing_param = sharing_param;
    // max_causal sizes the engine's exponent range.  computeTotalLikelihoodGivenConfigs does not bound the number of
    // causal SNPs of a row by maxCausalSNP (postcal.cpp:400-714): on that path size the store for the most a row may hold
    loc.max_causal = configsFile_.empty() ? MAX_causal : std::max(MAX_causal, (int)PIPSORT_KMAX);
    // PIPSORT_DEVICES=0,1,...: one process drives several GPUs -- the locus is replicated, the rank space (or the rows of
    // the explicit-configuration matrix) is split over the engines and the accumulator stores are added up (pipsort_merge)
    std::vector<int> devices;
    if (const char* dv = std::getenv("PIPSORT_DEVICES")) {
        int v = 0;
        bool have = false;
        for (const char* q = dv;; q++) {
            if (*q >= '0' && *q <= '9') { v = v * 10 + (*q - '0'); have = true; }
            else { if (have) devices.push_back(v); v = 0; have = false; if (!*q) break; }
        }
    }
    if (devices.empty()) devices.push_back(device);
    check(pipsort_create(&loc, devices[0], 0, &eng));
    for (size_t i = 1; i < devices.size(); i++) {
        pipsort_engine* x = nullptr;
        check(pipsort_create(&loc, devices[i], 0, &x));
        extra.push_back(x);
    }
}

PostCal::~PostCal() {
    for (pipsort_engine* x : extra) pipsort_destroy(x);
    pipsort_destroy(eng);
}

void PostCal::read_results() {
    pipsort_outputs out;
    out.total = &totalLikeLihoodLOG;
    out.postValues = postValues.data();
    out.noCausal = noCausal.data();
    out.sharedPips = sharedPips.data();
    out.sharedLL = sharedLL.data();
    out.notSharedLL = notSharedLL.data();
    check(pipsort_read_accumulators(eng, &out));
    check(pipsort_last_read_config_count(eng, &n_configs));
}

// postcal.cpp:716-1092: every union subset of size <= maxCausalSNP, every expansion
double PostCal::computeTotalLikelihood() {
    std::cout << "Max Causal = " << maxCausalSNP << std::endl;
    std::cout << "Union Snp Count = " << unionSnpCount << std::endl;
    if (maxCausalSNP > PIPSORT_KMAX) {
        std::cout << "Error: at most " << PIPSORT_KMAX << " causal SNPs are supported" << std::endl;
        std::exit(1);
    }
    uint64_t total = 0;
    check(pipsort_total_ranks(eng, maxCausalSNP, &total));
    if (extra.empty()) {
        check(pipsort_run_exhaustive(eng, maxCausalSNP, 0, total));
    } else {
        const int parts = 1 + (int)extra.size();
        std::vector<uint64_t> b(parts + 1);
        check(pipsort_shard_ranks(eng, maxCausalSNP, parts, b.data()));
        check(pipsort_run_exhaustive(eng, maxCausalSNP, b[0], b[1]));                      // asynchronous: all devices run
        for (int i = 1; i < parts; i++) check(pipsort_run_exhaustive(extra[i - 1], maxCausalSNP, b[i], b[i + 1]));
        for (pipsort_engine* x : extra) check(pipsort_merge(eng, x));
    }
    read_results();
    printf("num total configs = %llu\n", (unsigned long long)n_configs);
    return totalLikeLihoodLOG;
}

// postcal.cpp:400-714: the rows of the mmapped int16 matrix are the configurations (flags -b / -d / -e)
double PostCal::computeTotalLikelihoodGivenConfigs() {
    printf("Input configs given\n");
    printf("num total configs = %d\n", 0);
    std::cout << "Max Causal = " << maxCausalSNP << std::endl;
    std::cout << "Union Snp Count = " << unionSnpCount << std::endl;
    // util.cpp:26-49 safe_mmap_read_only + the size check of postcal.cpp:429-437
    const int fd = open(configsFile.c_str(), O_RDONLY);
    struct stat st;
    if (fd < 0) {
        printf("Could not open %s\n", configsFile.c_str());
        printf("mmap did not succeed\n");
        std::exit(1);
    }
    if (fstat(fd, &st) < 0) {
        printf("mmap did not succeed\n");
        std::exit(1);
    }
    const size_t len = (size_t)st.st_size;
    void* map = len ? mmap(nullptr, len, PROT_READ, MAP_SHARED, fd, 0) : nullptr;
    close(fd);
    if ((size_t)num_configs * (size_t)num_groups * sizeof(int16_t) != len || (len && map == MAP_FAILED)) {
        printf("config file is not the expected size\n");
        std::exit(1);
    }
    {
        const int parts = 1 + (int)extra.size();
        const int16_t* rows = static_cast<const int16_t*>(map);
        for (int i = 0; i < parts; i++) {
            const int64_t lo = (int64_t)num_configs * i / parts, hi = (int64_t)num_configs * (i + 1) / parts;
            check(pipsort_score_given_configs(i == 0 ? eng : extra[i - 1], rows + lo * num_groups, hi - lo, num_groups));
        }
        for (pipsort_engine* x : extra) check(pipsort_merge(eng, x));
    }
    if (len) munmap(map, len);
    read_results();
    printf("num total configs = %llu\n", (unsigned long long)n_configs);
    return totalLikeLihoodLOG;
}

// sss_postcal.cpp:20-48
std::vector<std::vector<int>> get_nbdplus(const std::vector<int>& cur, int unionSnpCount, int maxCausal) {
    std::vector<std::vector<int>> out;
    if ((int)cur.size() >= maxCausal) return out;
    std::vector<char> in(unionSnpCount, 0);
    for (int g : cur) in[g] = 1;
    for (int g = 0; g < unionSnpCount; g++) {
        if (in[g]) continue;
        std::vector<int> v;
        v.reserve(cur.size() + 1);
        v.push_back(g);
        v.insert(v.end(), cur.begin(), cur.end());
        std::sort(v.begin(), v.end());
        out.push_back(std::move(v));
    }
    return out;
}

// sss_postcal.cpp:50-69
std::vector<std::vector<int>> get_nbdminus(const std::vector<int>& cur) {
    std::vector<std::vector<int>> out;
    for (size_t drop = 0; drop < cur.size(); drop++) {
        std::vector<int> v;
        for (size_t j = 0; j < cur.size(); j++)
            if (j != drop) v.push_back(cur[j]);
        out.push_back(std::move(v));
    }
    return out;
}

// sss_postcal.cpp:72-99: added SNP outer (ascending), dropped SNP inner
std::vector<std::vector<int>> get_nbdzero(const std::vector<int>& cur, int unionSnpCount) {
    std::vector<std::vector<int>> out;
    std::vector<char> in(unionSnpCount, 0);
    for (int g : cur) in[g] = 1;
    const std::vector<std::vector<int>> minus = get_nbdminus(cur);
    for (int g = 0; g < unionSnpCount; g++) {
        if (in[g]) continue;
        for (const std::vector<int>& m : minus) {
            std::vector<int> v;
            v.reserve(m.size() + 1);
            v.push_back(g);
            v.insert(v.end(), m.begin(), m.end());
            std::sort(v.begin(), v.end());
            out.push_back(std::move(v));
        }
    }
    return out;
}

// sss_postcal.cpp:102-380.  The search loop (seed, neighbourhood order, hash map, sampling, stopping rules)
// stays on the host; every iteration scores the current configuration and all unseen neighbours with ONE
// batched engine call instead of the OpenMP loop of expand_and_compute_lkl.
double PostCal::sss_computeTotalLikelihood() {
    std::cout << "Max Causal = " << maxCausalSNP << std::endl;
    std::cout << "Union Snp Count = " << unionSnpCount << std::endl;
    if (maxCausalSNP > PIPSORT_KMAX) {
        std::cout << "Error: at most " << PIPSORT_KMAX << " causal SNPs are supported" << std::endl;
        std::exit(1);
    }
    const char* hl = std::getenv("PIPSORT_SSS_HOSTLOOP");
    if (!(hl && *hl == '1')) {
        // default: the whole search runs behind the C-ABI with the explored-configuration map in device memory
        int32_t iters = 0, why = 0;
        check(pipsort_sss(eng, maxCausalSNP, 1000, &iters, &why));       // total_iteration = 1000, sss_postcal.cpp:155
        if (why == 1) printf("hit break condition\n");
        if (why == 2) printf("hit convergence condition\n");
        sss_iterations = iters;
        read_results();
        return totalLikeLihoodLOG;
    }
    // PIPSORT_SSS_HOSTLOOP=1: the same search with the neighbourhood lists and the map on the host (std::map), one
    // batched scoring call per iteration -- kept as an independently written cross-check of pipsort_sss
    std::mt19937 gen(12345);                                             // sss_postcal.cpp:138
    std::vector<int> causal_locs;
    const int kmax = std::max(maxCausalSNP, 1);
    double old_sum_lkl = 0, sss_sum_lkl = 0;
    const int total_iteration = 1000;                                    // sss_postcal.cpp:155
    std::vector<int32_t> batch;
    std::vector<uint8_t> upd;
    std::vector<double> scored;
    int iter = 0;
    for (iter = 0; iter < total_iteration; iter++) {
        std::vector<std::vector<int>> nbd = get_nbdzero(causal_locs, unionSnpCount);
        const std::vector<std::vector<int>> nbdminus = get_nbdminus(causal_locs);
        const std::vector<std::vector<int>> nbdplus = get_nbdplus(causal_locs, unionSnpCount, maxCausalSNP);
        const size_t num_zero = nbd.size(), num_minus = nbdminus.size(), num_plus = nbdplus.size();
        nbd.insert(nbd.end(), nbdminus.begin(), nbdminus.end());
        nbd.insert(nbd.end(), nbdplus.begin(), nbdplus.end());

        // batch: [current configuration] + every neighbour the hash map has not seen
        const bool cur_updates = config_hashmap.find(causal_locs) == config_hashmap.end();   // :190-194
        std::vector<double> loglkls(nbd.size(), 0.0);
        std::vector<size_t> not_done;
        batch.assign((size_t)kmax, -1);
        for (size_t j = 0; j < causal_locs.size(); j++) batch[j] = causal_locs[j];
        upd.assign(1, cur_updates ? 1 : 0);
        for (size_t i = 0; i < nbd.size(); i++) {
            auto it = config_hashmap.find(nbd[i]);
            if (it != config_hashmap.end()) {
                loglkls[i] = it->second;
            } else {
                not_done.push_back(i);
                const size_t o = batch.size();
                batch.resize(o + kmax, -1);
                for (size_t j = 0; j < nbd[i].size(); j++) batch[o + j] = nbd[i][j];
                upd.push_back(1);
            }
        }
        scored.assign(upd.size(), 0.0);
        check(pipsort_score_union_configs(eng, batch.data(), (int64_t)upd.size(), kmax, upd.data(), scored.data()));
        for (size_t k = 0; k < not_done.size(); k++) loglkls[not_done[k]] = scored[k + 1];

        if (not_done.empty()) {                                          // :260-263
            printf("hit break condition\n");
            break;
        }
        if (iter >= 99) {                                                // the running sum is only consulted from here on
            pipsort_outputs o = {&sss_sum_lkl, nullptr, nullptr, nullptr, nullptr, nullptr};
            check(pipsort_read_accumulators(eng, &o));
        }
        if (iter >= 100 && (1 - std::exp(old_sum_lkl - sss_sum_lkl)) <= 0.001) {   // :265-270
            printf("hit convergence condition\n");
            break;
        }
        for (size_t i : not_done) config_hashmap[nbd[i]] = loglkls[i];   // :280-284

        // sampling, sss_postcal.cpp:289-343: one draw inside each group, then one across the groups
        double weight[3] = {0.0, 0.0, 0.0};
        size_t sample[3] = {nbd.size(), nbd.size(), nbd.size()};
        const size_t lo[3] = {0, num_zero, num_zero + num_minus};
        const size_t hi[3] = {num_zero, num_zero + num_minus, nbd.size()};
        (void)num_plus;
        for (int g = 0; g < 3; g++) {
            if (lo[g] == hi[g]) continue;
            const double max_log = *std::max_element(loglkls.begin() + lo[g], loglkls.begin() + hi[g]);
            std::vector<double> probs;
            for (size_t ii = lo[g]; ii < hi[g]; ii++) probs.push_back(std::exp(loglkls[ii] - max_log));
            std::discrete_distribution<size_t> dist(probs.begin(), probs.end());
            sample[g] = dist(gen);
            weight[g] = std::accumulate(probs.begin(), probs.end(), 0.0);
        }
        std::discrete_distribution<size_t> dist({weight[0], weight[1], weight[2]});
        const size_t grp = dist(gen);
        causal_locs = nbd[sample[grp] + lo[grp]];                        // :354
        old_sum_lkl = sss_sum_lkl;
    }
    sss_iterations = iter;
    read_results();
    return totalLikeLihoodLOG;
}

// postcal.cpp:1128-1243
std::vector<char> PostCal::findOptimalSetGreedy(std::vector<int>* rank, double inputRho, const std::string& outputFileName,
                                                double cutoff_threshold) {
    std::vector<char> causalSet(totalSnpCount, '0');
    const auto start = std::chrono::steady_clock::now();
    if (!configsFile.empty()) totalLikeLihoodLOG = computeTotalLikelihoodGivenConfigs();     // postcal.cpp:1134-1140
    else if (do_sss) totalLikeLihoodLOG = sss_computeTotalLikelihood();
    else totalLikeLihoodLOG = computeTotalLikelihood();
    const auto end = std::chrono::steady_clock::now();
    std::cout << "Time to eval all= " << std::chrono::duration_cast<std::chrono::microseconds>(end - start).count() << "[µs]"
              << std::endl;

    export2File(outputFileName + "_log.txt", std::exp(totalLikeLihoodLOG));
    double total_post = 0;
    for (int i = 0; i < totalSnpCount; i++) total_post = addlogSpace(total_post, postValues[i]);
    printf("\nTotal Likelihood = %e SNP=%d \n", total_post, totalSnpCount);
    total_post = totalLikeLihoodLOG;
    printf("total post as total likelihood log = %f\n", total_post);
    for (int i = 0; i < num_of_studies; i++) {
        printf("no causal just value %f\n", noCausal[i]);
        printf("Prob of no causal for study %d is %f\n", i, std::exp(noCausal[i] - total_post));
    }
    for (int i = 0; i < totalSnpCount; i++)
        if (special_exp(postValues[i], total_post) > 0.05) causalSet[i] = '1';           // postcal.cpp:1158-1163

    // per-study ranking by posterior (stdout diagnostics only, postcal.cpp:1168-1231)
    std::vector<ItemByProb> items;
    for (int i = 0; i < totalSnpCount; i++) items.push_back({std::exp(postValues[i] - total_post), i});
    printf("\n");
    int start_offset = 0, end_offset = 0;
    for (int s = 0; s < num_of_studies; s++) {
        end_offset += num_snps_all[s];
        printf("start offset = %d\n", start_offset);
        printf("end offset = %d\n", end_offset);
        std::sort(items.begin() + start_offset, items.begin() + end_offset,
                  [](const ItemByProb& a, const ItemByProb& b) { return a.p > b.p; });
        printf("sort complete %d\n", s);
        for (int i = 0; i < num_snps_all[s]; i++) (*rank)[start_offset + i] = items[i].idx;   // sic: postcal.cpp:1186-1188
        start_offset = end_offset;
    }
    std::cout << "threshold is " << cutoff_threshold << "\n";
    start_offset = end_offset = 0;
    for (int s = 0; s < num_of_studies; s++) {
        end_offset += num_snps_all[s];
        double rho = 0;
        int index = 0;
        while (rho < inputRho) {
            const double pr = special_exp(postValues[(*rank)[start_offset + index]], total_post);
            rho += pr;
            if (pr > cutoff_threshold) {
                const double pip = special_exp(postValues[start_offset + index], total_post);
                if (pip > 0.01) printf("%d %f\n", start_offset + index, pip);
            }
            index++;
            if (index >= num_snps_all[s]) break;
        }
        start_offset = end_offset;
    }
    printf("\n");
    return causalSet;
}

// postcal.h:288-336
void PostCal::printPost2File(const std::string& fileName) {
    const double total_post = totalLikeLihoodLOG;
    int start_offset = 0;
    for (int s = 0; s < num_of_studies; s++) {
        std::ofstream f((fileName + "_study" + std::to_string(s) + "_post.txt").c_str());
        f << "SNP_ID\tProb_in_pCausalSet" << std::endl;
        for (int j = 0; j < num_snps_all[s]; j++)
            f << (*SNP_NAME)[s][j] << "\t" << special_exp(postValues[start_offset + j], total_post) << std::endl;
        start_offset += num_snps_all[s];
    }
    {
        std::ofstream f((fileName + "_nocausal.txt").c_str());
        for (int s = 0; s < num_of_studies; s++) f << special_exp(noCausal[s], total_post) << std::endl;
    }
    {
        std::ofstream f((fileName + "_shared_pips.txt").c_str());
        f << "SNP_ID\tshared_pip\tshared_ll\tnotshared_ll" << std::endl;
        for (int i = 0; i < unionSnpCount; i++)
            f << all_snp_pos[i] << "\t" << special_exp(sharedPips[i], total_post) << "\t" << sharedLL[i] << "\t" << notSharedLL[i]
              << std::endl;
    }
}

}  // namespace pipsort_host
