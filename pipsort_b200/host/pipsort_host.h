// C++ host of pipsort_b200: the reference's Model / PostCal / command line re-written above the C-ABI
// (include/pipsort_b200.h).  Same flags, same input formats, same six output files (byte compatible:
// default ostream formatting = 6 significant digits); the hot path runs on the GPU engine only.
//
// Reference interfaces mirrored (paths relative to the PIPSORT source tree):
//   Model           model.h:22-321      (file reading, PSD fix, eigen-decomposition -> B, S')
//   PostCal         postcal.h:58-338    (findOptimalSetGreedy, printPost2File, SSS driver)
//   main            pipsort.cpp:68-228  (getopt string and its quirks)
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "pipsort_b200.h"

namespace pipsort_host {

// ---- util.cpp equivalents ---------------------------------------------------------------------------------
std::vector<std::string> read_dir(const std::string& fileName);              // pipsort.cpp:28-44
std::vector<int> read_sigma(const std::string& sample_size);                 // pipsort.cpp:46-66
void importData(const std::string& fileName, std::vector<double>& out);      // util.cpp:86-96
void importDataFirstColumn(const std::string& fileName, std::vector<std::string>& out);   // util.cpp:144-159
void importDataSecondColumn(const std::string& fileName, std::vector<double>& out);       // util.cpp:132-142
void importSnpMap(const std::string& file, int numCols, std::vector<std::string>& firstCol,
                  std::vector<std::vector<int>>& remainingCols);             // util.cpp:99-126
void export2File(const std::string& fileName, double data);                  // util.cpp:183-187 (append)

// ---- PostCal (postcal.h) ----------------------------------------------------------------------------------
class PostCal {
public:
    // what Model hands over (postcal.h:118): effective LD per study, z per study, K, the maps and the parameters
    PostCal(const std::vector<std::vector<double>>& sigma_eff, const std::vector<std::vector<double>>& z, double K,
            bool do_sss, int MAX_causal, const std::vector<std::vector<std::string>>* SNP_NAME, double sharing_param,
            double gamma, double t_squared, double s_squared, const std::vector<int>& sample_sizes,
            const std::vector<int>& num_snps_all, const std::vector<std::vector<int>>& idx_to_snp_map,
            const std::vector<std::string>& all_snp_pos, int device = 0, const std::string& configsFile = "",
            int num_configs = 0, int num_groups = 0);
    ~PostCal();
    PostCal(const PostCal&) = delete;
    PostCal& operator=(const PostCal&) = delete;

    double computeTotalLikelihood();          // postcal.cpp:716
    double sss_computeTotalLikelihood();      // sss_postcal.cpp:102
    double computeTotalLikelihoodGivenConfigs();   // postcal.cpp:400
    std::vector<char> findOptimalSetGreedy(std::vector<int>* rank, double inputRho, const std::string& outputFileName,
                                           double cutoff_threshold);          // postcal.cpp:1128
    void printPost2File(const std::string& fileName);                         // postcal.h:288

    // result members (postcal.h:62-99)
    std::vector<double> postValues, noCausal, sharedPips, sharedLL, notSharedLL;
    double totalLikeLihoodLOG = 0;
    int sss_iterations = 0;
    uint64_t n_configs = 0;

private:
    void read_results();
    pipsort_engine* eng = nullptr;                         // device 0 of the run: owns the merged accumulators
    std::vector<pipsort_engine*> extra;                    // PIPSORT_DEVICES: one more engine per additional GPU
    int num_of_studies, totalSnpCount, unionSnpCount, maxCausalSNP;
    bool do_sss;
    std::vector<int> num_snps_all;
    const std::vector<std::vector<std::string>>* SNP_NAME;
    std::vector<std::string> all_snp_pos;
    std::map<std::vector<int>, double> config_hashmap;     // postcal.h:98
    std::string configsFile;                               // postcal.h:84-86 (-b / -d / -e)
    int num_configs = 0, num_groups = 0;
};

// neighbourhoods of the stochastic shotgun search (sss_postcal.cpp:20-99)
std::vector<std::vector<int>> get_nbdplus(const std::vector<int>& causal_locs, int unionSnpCount, int maxCausal);
std::vector<std::vector<int>> get_nbdminus(const std::vector<int>& causal_locs);
std::vector<std::vector<int>> get_nbdzero(const std::vector<int>& causal_locs, int unionSnpCount);

// ---- Model (model.h) ----------------------------------------------------------------------------------------
class Model {
public:
    Model(const std::vector<std::string>& ldDir, const std::vector<std::string>& zDir, const std::string& snpMapFile,
          const std::string& configsFile, int num_configs, int num_groups, bool do_sss, const std::vector<int>& sample_sizes, const std::string& outputFileName, int totalCausalSNP,
          double sharing_param, double rho, double gamma, double tau_sqr, double sigma_g_squared, double cutoff_threshold,
          int device = 0);
    ~Model();
    void run();
    void finishUp();
    PostCal* post = nullptr;

private:
    int num_of_studies;
    double rho, cutoff_threshold;
    std::string outputFileName;
    std::vector<int> num_snps_all;
    std::vector<std::vector<std::string>> snpNames;
    std::vector<char> pcausalSet;
    std::vector<int> rank;
};

int pipsort_main(int argc, char* argv[]);

}  // namespace pipsort_host
