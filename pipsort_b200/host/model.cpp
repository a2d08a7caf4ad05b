// Model (model.h:22-321): read the locus, make it positive definite, eigen-decompose, build PostCal, write
// the per-study set files.  Host code; the O(n^3) pre-processing runs once per locus.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "pipsort_host.h"

namespace pipsort_host {

Model::Model(const std::vector<std::string>& ldDir, const std::vector<std::string>& zDir, const std::string& snpMapFile,
             const std::string& configsFile, int num_configs, int num_groups, bool do_sss, const std::vector<int>& sample_sizes, const std::string& outName, int totalCausalSNP,
             double sharing_param, double rho_, double gamma, double tau_sqr, double sigma_g_squared, double cutoff, int device)
    : num_of_studies((int)ldDir.size()), rho(rho_), cutoff_threshold(cutoff), outputFileName(outName) {
    std::vector<std::vector<double>> sigma, z_score;
    std::vector<std::vector<int>> idx_to_snp_map(num_of_studies), idx_to_union_pos_map(num_of_studies);
    std::vector<std::string> all_snp_pos;
    for (int i = 0; i < num_of_studies; i++) {
        std::vector<double> ld, z;
        std::vector<std::string> names;
        importData(ldDir[i], ld);
        importDataFirstColumn(zDir[i], names);
        importDataSecondColumn(zDir[i], z);
        const int numSnps = (int)std::sqrt((double)ld.size());                                   // model.h:98
        num_snps_all.push_back(numSnps);
        if (numSnps != (int)names.size()) {
            printf("ERROR: LD matrix is size %d x %d but zscores has %lu snps\n. Check LD file for nans.\n", numSnps, numSnps,
                   (unsigned long)names.size());
            std::exit(1);
        }
        printf("pushing back num snps %d for study %d\n", i, numSnps);                           // sic: model.h:104
        ld.resize((size_t)numSnps * numSnps);
        sigma.push_back(std::move(ld));
        snpNames.push_back(std::move(names));
        z_score.push_back(std::move(z));
    }
    importSnpMap(snpMapFile, num_of_studies + 1, all_snp_pos, idx_to_snp_map);
    for (int i = 0; i < num_of_studies; i++) {
        for (size_t j = 0; j < idx_to_snp_map[i].size(); j++)
            if (idx_to_snp_map[i][j] >= 0) idx_to_union_pos_map[i].push_back((int)j);
        if ((int)idx_to_union_pos_map[i].size() != num_snps_all[i]) {                            // model.h:140-143
            printf("Invariant does not hold\n");
            std::exit(1);
        }
    }
    int total = 0;
    for (int n : num_snps_all) total += n;
    pcausalSet.assign(total, '0');
    rank.assign(total, 0);
    double K = 0.0;
    for (int i = 0; i < num_of_studies; i++) {
        const auto t0 = std::chrono::steady_clock::now();
        // model.h:171-264 on the GPU (cuSOLVER LU / eigensolver + the engine's kernels): there is no host compute path
        pipsort_prep_info info;
        {
            std::vector<double> eff(sigma[i].size());
            if (pipsort_preprocess_study(device, num_snps_all[i], sigma[i].data(), z_score[i].data(), eff.data(), &info) != 0) {
                std::cout << pipsort_last_error() << std::endl;
                std::exit(1);
            }
            sigma[i].swap(eff);
        }
        const auto t1 = std::chrono::steady_clock::now();
        std::cout << "Time to make psd + eigen decomp = " << std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count()
                  << "[µs] (diagonal shift " << info.add_diag << ", smallest eigenvalue " << info.min_abs_eig << ")" << std::endl;
        K += info.K;
    }
    post = new PostCal(sigma, z_score, K, do_sss, totalCausalSNP, &snpNames, sharing_param, gamma, tau_sqr, sigma_g_squared,
                       sample_sizes, num_snps_all, idx_to_snp_map, all_snp_pos, device, configsFile, num_configs, num_groups);
}

Model::~Model() { delete post; }

void Model::run() { pcausalSet = post->findOptimalSetGreedy(&rank, rho, outputFileName, cutoff_threshold); }   // model.h:273

// model.h:282-310
void Model::finishUp() {
    int start = 0;
    for (int s = 0; s < num_of_studies; s++) {
        std::ofstream f((outputFileName + "_study" + std::to_string(s) + "_set.txt").c_str());
        for (int j = 0; j < num_snps_all[s]; j++)
            if (pcausalSet[start + j] == '1') f << snpNames[s][j] << std::endl;
        start += num_snps_all[s];
    }
    post->printPost2File(outputFileName);
}

}  // namespace pipsort_host
