// host_capi.cpp — C entry points over the host pipeline, for ctypes / other FFIs.
#include "zwz_host.hpp"

#include <cstring>
#include <exception>
#include <iostream>

using namespace zwzhost;
namespace zwzhost {
std::string md5_of_file_on(int device, const std::string &file_path);
}

extern "C" {

int zwz_host_compress(const char *input_dir, const char *output_dir, int world_rank, int world_size, int device, int level) {
    try {
        RunConfig &c = config();
        c.world_rank = world_rank;
        c.world_size = world_size;
        c.device = device;
        c.level = level;
        stats() = RunStats();
        std::string record = (std::filesystem::path(input_dir).parent_path() / "sorted_files_by_size.txt").string();
        if (world_rank == 0) record = sort_files_by_size(input_dir);
        if (world_rank < count_non_empty_lines(record)) do_compression(input_dir, output_dir, record, world_rank);
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "zwz_host_compress: " << e.what() << std::endl;
        return -1;
    }
}

int zwz_host_decompress(const char *input_dir, const char *output_dir, int device) {
    try {
        config().device = device;
        stats() = RunStats();
        do_decompression(input_dir, output_dir);
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "zwz_host_decompress: " << e.what() << std::endl;
        return -1;
    }
}

int zwz_host_md5_of_file(const char *path, int device, char hex_out[33]) {
    try {
        std::string h = md5_of_file_on(device, path);
        std::memset(hex_out, 0, 33);
        std::memcpy(hex_out, h.data(), h.size() < 32 ? h.size() : 32);
        return h.empty() ? 1 : 0;
    } catch (const std::exception &e) {
        std::cerr << "zwz_host_md5_of_file: " << e.what() << std::endl;
        return -1;
    }
}

int zwz_host_sort_files_by_size(const char *dir, char *record_path_out, size_t cap) {
    try {
        std::string r = sort_files_by_size(dir);
        if (r.size() + 1 > cap) return -1;
        std::memcpy(record_path_out, r.c_str(), r.size() + 1);
        return 0;
    } catch (const std::exception &e) {
        return -1;
    }
}

void zwz_host_last_stats(uint64_t out[6]) {
    const RunStats &s = stats();
    out[0] = s.files;
    out[1] = s.records;
    out[2] = s.raw_bytes;
    out[3] = s.payload_bytes;
    out[4] = s.md5_match;
    out[5] = s.md5_mismatch;
}
}
