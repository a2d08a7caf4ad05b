// File readers / writers of the host side (no numerics: the pre-processing runs on the GPU, csrc/prep.cuh).
// Behaviour follows util.cpp / pipsort.cpp of the reference (cited per function); the code is new.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "pipsort_host.h"

namespace pipsort_host {

// pipsort.cpp:28-44: one path per non-empty line
std::vector<std::string> read_dir(const std::string& fileName) {
    std::ifstream fin(fileName.c_str());
    if (!fin) {
        std::cout << "Error: unable to open " << fileName << std::endl;
        std::exit(1);
    }
    std::vector<std::string> out;
    std::string line;
    while (fin.good()) {
        std::getline(fin, line);
        if (!line.empty()) out.push_back(line);
    }
    return out;
}

// pipsort.cpp:46-66: comma separated non-negative integers, anything else is a format error
std::vector<int> read_sigma(const std::string& text) {
    std::vector<int> sizes;
    std::string cur;
    for (char ch : text) {
        if (ch == ',') {
            sizes.push_back((int)std::stod(cur));
            cur.clear();
        } else if (std::isdigit((unsigned char)ch)) {
            cur += ch;
        } else {
            std::cout << "Error: sample size is not in the right format" << std::endl;
            std::exit(1);
        }
    }
    if (!cur.empty()) sizes.push_back((int)std::stod(cur));
    return sizes;
}

// util.cpp:86-96: whitespace separated doubles; extraction stops at the first token that is not a number
void importData(const std::string& fileName, std::vector<double>& out) {
    std::ifstream file(fileName.c_str());
    if (!file) {
        std::cout << "Unable to open file; This is why";
        std::exit(1);
    }
    double v;
    while (file >> v) out.push_back(v);
}

// util.cpp:144-159
void importDataFirstColumn(const std::string& fileName, std::vector<std::string>& out) {
    std::ifstream fin(fileName.c_str());
    std::string line, tok;
    while (std::getline(fin, line)) {
        std::istringstream iss(line);
        iss >> tok;                      // an empty line repeats the previous token, like the reference
        out.push_back(tok);
    }
}

// util.cpp:132-142
void importDataSecondColumn(const std::string& fileName, std::vector<double>& out) {
    std::ifstream fin(fileName.c_str());
    std::string line, name;
    double v = 0.0;
    while (std::getline(fin, line)) {
        std::istringstream iss(line);
        iss >> name;
        iss >> v;
        out.push_back(v);
    }
}

// util.cpp:99-126: rsid,idx0,idx1 per line
void importSnpMap(const std::string& file, int numCols, std::vector<std::string>& firstCol,
                  std::vector<std::vector<int>>& remainingCols) {
    std::ifstream f(file.c_str());
    if (!f.is_open()) {
        std::cout << "Could not open file\n";
        std::exit(1);
    }
    std::string line, word;
    while (std::getline(f, line)) {
        std::stringstream s(line);
        for (int i = 0; i < numCols; i++) {
            std::getline(s, word, ',');
            if (i == 0) firstCol.push_back(word);
            else remainingCols[i - 1].push_back(std::stoi(word));
        }
    }
}

// util.cpp:183-187: appended, default stream formatting
void export2File(const std::string& fileName, double data) {
    std::ofstream outfile(fileName.c_str(), std::ios::out | std::ios::app);
    outfile << data << std::endl;
}

}  // namespace pipsort_host
