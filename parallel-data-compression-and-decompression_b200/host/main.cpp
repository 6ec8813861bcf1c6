// main.cpp — `main compress|decompress <source directory> <output directory>` (same CLI as the reference's main.cpp:78-159).
//
// The reference is an MPI program; its ranks exchange nothing but the path of the record file (main.cpp:27,35). Here a
// "rank" is one process driving one B200. Rank and size come from the launcher's environment (ZWZ_RANK/ZWZ_WORLD, or the
// Open MPI / PMI / torchrun variables); with `ZWZ_GPUS=N` and no launcher, this process forks N-1 siblings itself so that
// `ZWZ_GPUS=8 main compress src dst` shards one directory over the 8 GPUs of a box. Every rank computes the same
// size-descending deal on its own (same walk, same comparator, same libstdc++ sort), so no broadcast is needed; rank 0
// also writes <src>/../sorted_files_by_size.txt like the reference does.
#include "zwz_host.hpp"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

using namespace zwzhost;

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void remove_trailing_slash(std::string &path) { // main.cpp:72-76
    if (!path.empty() && path.back() == '/') path.pop_back();
}

// every rank waits here until rank 0 has published the record file for THIS run (the reference uses MPI_Bcast + MPI_Barrier)
static bool wait_for_file(const std::string &path, double not_before, double timeout_s) {
    double t0 = now_s();
    while (now_s() - t0 < timeout_s) {
        struct stat st {};
        if (stat(path.c_str(), &st) == 0 && (double) st.st_mtime + 1.0 >= not_before) return true;
        usleep(2000);
    }
    return false;
}

static void compress(const std::string &folder_path, const std::string &output_path) { // main.cpp:10-54
    const RunConfig &cfg = config();
    std::string file_record = (std::filesystem::path(folder_path).parent_path() / "sorted_files_by_size.txt").string();
    if (cfg.world_rank == 0) {
        std::cout << "Compressing folder: " << folder_path << std::endl;
        std::string tmp_record = sort_files_by_size(folder_path);
        std::cout << "File record saved location: " << tmp_record << std::endl;
        file_record = tmp_record;
        std::ofstream(output_path + "/.zwz_record_ready") << file_record << "\n";
    } else {
        // the marker lives in the (fresh) output directory, so a stale record file from an earlier run is never trusted
        if (!wait_for_file(output_path + "/.zwz_record_ready", 0.0, 600.0)) {
            std::cerr << "Rank: " << cfg.world_rank << " - timed out waiting for the file record" << std::endl;
            return;
        }
    }
    std::cout << "file_record: " << file_record << std::endl;
    timing_mark("file record ready");
    int file_count = count_non_empty_lines(file_record);
    if (cfg.world_rank < file_count) {
        do_compression(folder_path, output_path, file_record, cfg.world_rank);
    } else {
        std::cout << "Rank: " << cfg.world_rank << " - No file to compress" << std::endl;
    }
    std::cout << "main - Rank: " << cfg.world_rank << " - do_compression finished" << std::endl;
}

static void decompress(const std::string &source_path, const std::string &output_path) { // main.cpp:56-70
    const RunConfig &cfg = config();
    if (cfg.world_rank == 0) {
        if (cfg.world_size > 1) {
            std::cout << "Decompression is not supported in MPI parallel mode.\n";
            std::cout << "Only use one process to decompress.\n";
        }
        do_decompression(source_path, output_path);
    }
}

namespace {

const char *kReadyMarker = "/.zwz_record_ready";

// rank 0 only (main.cpp:104-129): the source must exist; the output directory is created when missing
bool prepare_paths(const std::string &source_path, const std::string &output_path) {
    namespace fs = std::filesystem;
    std::error_code ec;
    if (!fs::exists(source_path, ec)) {
        std::cerr << "Source path does not exist.\n";
        return false;
    }
    if (!fs::exists(output_path, ec)) {
        if (mkdir(output_path.c_str(), 0777) == -1) {
            perror("Failed to create output directory");
            return false;
        }
        return true;
    }
    if (!fs::is_directory(output_path, ec)) {
        std::cerr << "Output path is not a directory.\n";
        return false;
    }
    return true;
}

// ZWZ_GPUS=N without an external launcher: ranks 1..N-1 are forked here, one process per GPU, before this process has
// made any CUDA call. Returns the children (empty in a child).
std::vector<pid_t> fork_ranks(int n) {
    RunConfig &cfg = config();
    std::vector<pid_t> kids;
    cfg.world_size = n;
    for (int r = 1; r < n; ++r) {
        pid_t pid = fork();
        if (pid == 0) {
            cfg.world_rank = r;
            const int ndev = visible_gpu_count();
            cfg.device = ndev > 0 ? r % ndev : 0;
            return {};
        }
        kids.push_back(pid);
    }
    return kids;
}

void print_summary(const std::string &operation, double seconds) { // main.cpp:148-156
    std::cout << "========================================\n"
              << "Operation: " << operation << '\n'
              << "Processor Count: " << config().world_size << '\n'
              << "Time Taken: " << seconds << " seconds\n"
              << "========================================\n";
}

} // namespace

int main(int argc, char *argv[]) {
    const double start_time = now_s();
    config_from_env();
    RunConfig &cfg = config();
    timing_mark("configured");

    if (argc < 4) { // main.cpp:88-92
        std::cerr << "Usage: " << argv[0] << " <compress/decompress> <source directory path> <output directory path>\n";
        return 1;
    }
    const std::string operation = argv[1];
    std::string source_path = argv[2], output_path = argv[3];
    remove_trailing_slash(source_path);
    remove_trailing_slash(output_path);
    std::cout << "source_path: " << source_path << '\n';
    std::cout << "output_path: " << output_path << '\n';

    const bool compressing = operation == "compress";
    if (cfg.world_rank == 0) {
        if (!prepare_paths(source_path, output_path)) return 1;
        if (compressing) std::remove((output_path + kReadyMarker).c_str());
    }
    const char *g = std::getenv("ZWZ_GPUS");
    const int self_gpus = (g && cfg.world_size == 1) ? std::atoi(g) : 0;
    std::vector<pid_t> kids;
    if (self_gpus > 1 && compressing) kids = fork_ranks(self_gpus);

    int rc = 0;
    try {
        if (compressing) {
            compress(source_path, output_path);
        } else if (operation == "decompress") {
            decompress(source_path, output_path);
        } else {
            std::cerr << "Invalid operation: " << operation << ". Please use 'compress' or 'decompress'.\n";
            return 1;
        }
    } catch (const std::exception &e) {
        std::cerr << "Rank: " << cfg.world_rank << " - fatal: " << e.what() << std::endl;
        rc = 2;
    }
    timing_mark("operation finished");
    if (cfg.strict && stats().bad_records != 0 && rc == 0) {
        std::cerr << "ZWZ_STRICT: " << stats().bad_records << " record(s) did not decode cleanly" << std::endl;
        rc = 4;
    }
    if (self_gpus > 1 && compressing && cfg.world_rank != 0) _exit(rc); // a forked rank: no summary, no waiting
    for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        if ((!WIFEXITED(st) || WEXITSTATUS(st) != 0) && rc == 0) rc = 3;
    }
    if (cfg.world_rank == 0) {
        if (compressing) std::remove((output_path + kReadyMarker).c_str());
        print_summary(operation, now_s() - start_time);
    }
    return rc;
}
