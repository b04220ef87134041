// Forwarding header for opq/src/common.h: the helpers its callers use (get_sort_results,
// get_vector_of_strings_from_file_lines, get_base_name, Init2DArray/Delete2DArray).
#pragma once
#include <fstream>
#include <string>
#include <vector>
#include "../ivfopq_gpu.hpp"
static inline void get_vector_of_strings_from_file_lines(const std::string file_name, std::vector<std::string>& out) {
    std::ifstream in(file_name.c_str());
    std::string line;
    out.clear();
    while (std::getline(in, line)) out.push_back(line);
}
static inline std::string get_base_name(const std::string path) {
    const size_t slash = path.find_last_of("/\\");
    const std::string file = path.substr(slash + 1);
    return file.substr(0, file.find_last_of('.'));
}
