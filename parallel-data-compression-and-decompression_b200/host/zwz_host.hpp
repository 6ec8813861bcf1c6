// zwz_host.hpp — host side of the shard-based DEFLATE pipeline, mirroring the reference's own interface (process.hpp:12-42)
// so that a maintainer can swap files one for one. Everything that touches bytes goes through the C ABI in
// include/zwz_cuda.h (CUDA kernels); this layer only walks directories, deals files to ranks, reads/writes files and
// serialises `.zwz` records byte-compatibly with compression.cpp:73-104 / decompression.cpp:65-92.
#pragma once
#include <cstddef>
#include <cstdint>
#include <filesystem>
#include <string>
#include <vector>

#include "zwz_cuda.h"

namespace zwzhost {

constexpr std::size_t CHUNK_SIZE = ZWZ_CHUNK_SIZE; // process.hpp:12
constexpr std::size_t MD5_DATA_SIZE = 32;          // process.hpp:14

struct FileEntry { // process.hpp:16-19
    std::string relpath;
    off_t size;
};

// ---- same names and meaning as process.hpp:37-42 ----
std::string sort_files_by_size(const std::filesystem::path &path);                 // file_sort.cpp:24-43
int count_non_empty_lines(const std::string &file_path);                           // file_tools.cpp:6-23
void do_compression(const std::string &input_dir, const std::string &output_dir, const std::string &file_record, int world_rank);
void do_decompression(const std::string &input_dir, const std::string &output_dir);
std::string md5_of_file(const std::string &file_path);                             // verification.cpp:6-30
bool is_md5_match(const std::string &file_path, const std::string &expected_md5);  // verification.cpp:32-36

// ---- run configuration (the reference takes these from MPI; here: environment or the launcher) ----
struct RunConfig {
    int world_rank = 0;
    int world_size = 1;
    int device = 0;           // CUDA device this rank drives
    int level = 0;            // 0 = library default
    bool verbose = false;     // per-file lines like the reference prints (compression.cpp:100, decompression.cpp:90,145)
    bool verify_all = false;  // also print an MD5 verdict for files whose last record arrived out of order
    bool strict = false;      // ZWZ_STRICT=1: name every record that did not decode to a clean end of stream, exit code 4
                              // (the reference ignores zlib's return codes, decompression.cpp:31 — and so does the default)
    std::string run_id;       // ZWZ_RUN_ID or the nonce of a self-launch: keys the files the ranks of one run leave for each other
    bool quarantine = false;  // ZWZ_QUARANTINE=1: move files whose MD5 does not match to <output dir>/bad/ (README.md:175,186 promises
                              // it; the reference's code leaves them in place, and so does the default)
    std::size_t batch_bytes = (std::size_t) 128 << 20; // per worker (ZWZ_BATCH_MB)
};
RunConfig &config();
void config_from_env();
int visible_gpu_count(); // without initialising CUDA when possible
void warm_device();      // bring up this rank's CUDA context (errors are left for the first real use to report)

// deterministic walk + size-descending sort, shared by sort_files_by_size and by ranks that recompute the deal
std::vector<FileEntry> collect_and_sort(const std::filesystem::path &path);

struct RunStats {
    uint64_t files = 0, records = 0, raw_bytes = 0, payload_bytes = 0, md5_match = 0, md5_mismatch = 0, bad_records = 0;
    double t_read = 0, t_gpu = 0, t_write = 0, t_init = 0; // seconds, printed with ZWZ_TIMING=1
};
double now_seconds();
void print_timing(const char *what);
void timing_mark(const char *what);
int timing_level();
RunStats &stats();

} // namespace zwzhost

// C entry points for ctypes / other FFIs (same semantics as the C++ functions above)
extern "C" {
// NOT re-entrant: these entry points set the process-wide run configuration and statistics (one run per process at a time, like
// the reference's `main`). zwz_host_compress with world_rank != 0 reads <input_dir>/../sorted_files_by_size.txt: the caller runs
// rank 0's zwz_host_sort_files_by_size (or its whole zwz_host_compress) first — the record file is published by rename, so a
// reader never sees half of it, but only the caller can know that it is THIS run's.
int zwz_host_compress(const char *input_dir, const char *output_dir, int world_rank, int world_size, int device, int level);
int zwz_host_decompress(const char *input_dir, const char *output_dir, int device);
int zwz_host_md5_of_file(const char *path, int device, char hex_out[33]);
int zwz_host_sort_files_by_size(const char *dir, char *record_path_out, size_t cap);
void zwz_host_last_stats(uint64_t out[6]);
}
