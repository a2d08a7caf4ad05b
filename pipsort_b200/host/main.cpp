// PIPSORT command line (pipsort.cpp:68-228) in front of the GPU engine.  Flag string, defaults and quirks are
// the reference's: `-m` falls through into `-n` (so -n must come after -m), flags without an argument exit 1,
// -r / -a / -k / -f only influence stdout; -b/-d/-e select the explicit-configuration path (postcal.cpp:400-714).
// Not supported (exit 1 with a message): more or fewer than two studies.
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "pipsort_host.h"

namespace pipsort_host {

int pipsort_main(int argc, char* argv[]) {
    int totalCausalSNP = 3;                      // pipsort.cpp:69-77
    double gamma = 0.01, sharing_param = 0.75, rho = 0.95, tau_sqr = 0.52, sigma_g_squared = 5.2, cutoff_threshold = 0;
    std::string ldFile, zFile, snpMapFile, outputFileName, sample_s, num_causal_s, configsFile;
    int num_groups = 0, num_configs = 0, sss_flag = 0, oc = 0, device = 0;
    if (const char* dv = std::getenv("PIPSORT_DEVICE")) device = std::atoi(dv);

    while ((oc = getopt(argc, argv, "vhl:o:z:m:p:r:c:k:g:f:t:s:n:a:b:d:e:q:x")) != -1) {
        if (optarg == NULL || *optarg == '\0') {  // pipsort.cpp:92-95
            printf("optarg is NULL\n");
            std::exit(1);
        }
        switch (oc) {
            case 'l': ldFile = optarg; break;
            case 'o': outputFileName = optarg; break;
            case 'z': zFile = optarg; break;
            case 'm': snpMapFile = optarg;       // falls through on purpose: pipsort.cpp:128-132 has no break here
                [[fallthrough]];
            case 'n': sample_s = optarg; break;
            case 'b': configsFile = optarg; break;
            case 'd': num_configs = std::atoi(optarg); break;
            case 'e': num_groups = std::atoi(optarg); break;
            case 'p': sharing_param = std::atof(optarg); break;
            case 'r': rho = std::atof(optarg); break;
            case 'c': totalCausalSNP = std::atoi(optarg); break;
            case 'k': num_causal_s = optarg; break;
            case 'g': gamma = std::atof(optarg); break;
            case 'f': break;
            case 't': tau_sqr = std::atof(optarg); break;
            case 's': sigma_g_squared = std::atof(optarg); break;
            case 'q': sss_flag = std::stoi(optarg); break;
            case ':':
            case '?':
            case 'a': cutoff_threshold = std::atof(optarg); break;
            default: break;
        }
    }
    if (ldFile.empty() || zFile.empty() || snpMapFile.empty() || outputFileName.empty() || sample_s.empty()) {
        std::cout << "Error: -l, -z, -o, and -n are required" << std::endl;
        std::exit(1);
    }
    if (!configsFile.empty()) {                  // pipsort.cpp:189-197 (no exit after the num_groups message: sic)
        if (num_configs <= 0) {
            std::cout << "Number of configs must be greater than 0" << std::endl;
            std::exit(1);
        }
        if (num_groups <= 0) std::cout << "Number of groups must be greater than 0" << std::endl;
    }
    const std::vector<std::string> ldDir = read_dir(ldFile), zDir = read_dir(zFile);
    const std::vector<int> sample_sizes = read_sigma(sample_s);
    if (ldDir.size() != zDir.size() || ldDir.size() != sample_sizes.size()) {
        std::cout << "Error: LD files, Z files, and sample sizes do not match in number" << std::endl;
        std::exit(1);
    }
    Model m(ldDir, zDir, snpMapFile, configsFile, num_configs, num_groups, sss_flag == 1, sample_sizes, outputFileName, totalCausalSNP, sharing_param, rho, gamma,
            tau_sqr, sigma_g_squared, cutoff_threshold, device);
    m.run();
    m.finishUp();
    return 0;
}

}  // namespace pipsort_host

int main(int argc, char* argv[]) { return pipsort_host::pipsort_main(argc, argv); }
