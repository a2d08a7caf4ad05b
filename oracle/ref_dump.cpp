// TEST INFRASTRUCTURE ONLY -- full-precision dump harness around the UNMODIFIED reference.
//
// The reference prints every number with 6 significant digits (postcal.h:288-336).  The parity
// gates need 1e-8 / 1e-10, so this harness drives the reference's own Model / PostCal classes
// (compiled from the sources where they lie under /root/reference, see oracle/Makefile) exactly
// the way /root/reference/pipsort.cpp:204-226 does, and then writes the private log-space arrays
// (postcal.h:62-99) with 17 significant digits to <out>_raw.txt.  It contains no algorithm of its
// own; it is linked only into oracle/_ref/pipsort_ref_dump and never into the product.
#include <armadillo>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <numeric>
#include <string>
#include <unistd.h>
#include <vector>

#include "util.h"
#define private public
#include "postcal.h"
#undef private
#include "model.h"

// defined in the reference's pipsort.cpp, which is compiled with -Dmain=pipsort_reference_main
vector<string> read_dir(string fileName);
vector<int> read_sigma(string sample_size);

int main(int argc, char* argv[]) {
    int totalCausalSNP = 3;  // defaults: pipsort.cpp:69-77
    double gamma = 0.01, sharing_param = 0.75, rho = 0.95, tau_sqr = 0.52, sigma_g_squared = 5.2, cutoff = 0;
    string ldFile, zFile, snpMapFile, out, sample_s, configsFile;
    int num_groups = 0, num_configs = 0, sss_flag = 0, oc;
    while ((oc = getopt(argc, argv, "l:o:z:m:n:p:c:g:t:s:q:b:d:e:")) != -1) {
        switch (oc) {
            case 'l': ldFile = optarg; break;
            case 'o': out = optarg; break;
            case 'z': zFile = optarg; break;
            case 'm': snpMapFile = optarg; break;
            case 'n': sample_s = optarg; break;
            case 'p': sharing_param = atof(optarg); break;
            case 'c': totalCausalSNP = atoi(optarg); break;
            case 'g': gamma = atof(optarg); break;
            case 't': tau_sqr = atof(optarg); break;
            case 's': sigma_g_squared = atof(optarg); break;
            case 'q': sss_flag = atoi(optarg); break;
            case 'b': configsFile = optarg; break;
            case 'd': num_configs = atoi(optarg); break;
            case 'e': num_groups = atoi(optarg); break;
            default: return 2;
        }
    }
    if (ldFile.empty() || zFile.empty() || snpMapFile.empty() || out.empty() || sample_s.empty()) {
        fprintf(stderr, "usage: pipsort_ref_dump -l LD -z Z -m MAP -n N0,N1 -o OUT [-c -p -g -t -s -q -b -d -e]\n");
        return 2;
    }
    vector<string> ldDir = read_dir(ldFile), zDir = read_dir(zFile);
    vector<int> sample_sizes = read_sigma(sample_s);
    vector<int> num_causal = read_sigma("");
    omp_set_num_threads(1);
    Model m(ldDir, zDir, snpMapFile, configsFile, num_configs, num_groups, sss_flag == 1, sample_sizes, num_causal, out,
            totalCausalSNP, sharing_param, rho, false, gamma, tau_sqr, sigma_g_squared, cutoff);
    m.run();
    m.finishUp();

    PostCal* pc = m.post;
    FILE* f = fopen((out + "_raw.txt").c_str(), "w");
    if (!f) return 3;
    int N = pc->totalSnpCount, U = pc->unionSnpCount, S = pc->num_of_studies;
    fprintf(f, "total %.17g\n", pc->totalLikeLihoodLOG);
    fprintf(f, "K %.17g\n", arma::as_scalar(pc->statMatrixtTran * pc->statMatrix));
    fprintf(f, "dims %d %d %d\n", S, N, U);
    for (int s = 0; s < S; s++) fprintf(f, "noCausal %d %.17g\n", s, pc->noCausal[s]);
    for (int i = 0; i < N; i++) fprintf(f, "postValues %d %.17g\n", i, pc->postValues[i]);
    for (int g = 0; g < U; g++)
        fprintf(f, "shared %d %.17g %.17g %.17g\n", g, pc->sharedPips[g], pc->sharedLL[g], pc->notSharedLL[g]);
    fclose(f);
    return 0;
}
