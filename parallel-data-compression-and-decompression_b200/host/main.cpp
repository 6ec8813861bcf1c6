// main.cpp — `main compress|decompress <source directory> <output directory>` (same CLI as the reference's main.cpp:78-159).
//
// The reference is an MPI program; its ranks exchange nothing but the path of the record file (main.cpp:27,35). Here a
// "rank" is one process driving one B200. Rank and size come from the launcher's environment (ZWZ_RANK/ZWZ_WORLD, or the
// Open MPI / PMI / torchrun variables); with `ZWZ_GPUS=N` and no launcher, this process forks N-1 siblings itself so that
// `ZWZ_GPUS=8 main compress src dst` shards one directory over the 8 GPUs of a box. Every rank computes the same
// size-descending deal on its own (same walk, same comparator, same libstdc++ sort), so no broadcast is needed; rank 0
// also writes <src>/../sorted_files_by_size.txt like the reference does.
//
// Where the reference broadcasts the record path and passes a barrier (main.cpp:27-41), the ranks here meet at a marker file
// in the output directory that is keyed to the RUN: it carries a run id — a nonce the forking parent made (ZWZ_GPUS), or
// ZWZ_RUN_ID from an external launcher — and a rank only accepts a marker with its own run id (without any id: one no older
// than two minutes). A marker left by a killed run is therefore never trusted, rank 0 never has to remove it behind the
// others' backs, and a rank that does not see it within ten minutes fails with a non-zero exit code.
#include "zwz_host.hpp"

#include <chrono>
#include <cstdlib>
#include <algorithm>
#include <cerrno>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sys/stat.h>
#include <thread>
#include <sys/wait.h>
#include <unistd.h>

using namespace zwzhost;

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void remove_trailing_slash(std::string &path) { // main.cpp:72-76
    if (!path.empty() && path.back() == '/') path.pop_back();
}

static std::string g_run_id; // ZWZ_RUN_ID, or the nonce of the self-launch

static const char *kReadyMarker = "/.zwz_record_ready";

// marker: "<run id>\n<record file>\n"
static void publish_marker(const std::string &output_path, const std::string &file_record) {
    const std::string tmp = output_path + kReadyMarker + ".tmp";
    {
        std::ofstream f(tmp);
        f << g_run_id << "\n" << file_record << "\n";
    }
    std::rename(tmp.c_str(), (output_path + kReadyMarker).c_str()); // atomically: a reader sees all of it or none
}

// every rank but 0 waits here until rank 0 has published the record file for THIS run (the reference uses MPI_Bcast + MPI_Barrier)
static bool wait_for_marker(const std::string &output_path, double started, double timeout_s, std::string &file_record) {
    const std::string path = output_path + kReadyMarker;
    double t0 = now_s();
    while (now_s() - t0 < timeout_s) {
        std::ifstream f(path);
        std::string id, rec;
        if (f.is_open() && std::getline(f, id) && std::getline(f, rec)) {
            bool ours;
            if (!g_run_id.empty()) {
                ours = id == g_run_id;
            } else { // no run id to compare: accept a marker written around the time this rank started
                struct stat st {};
                ours = id.empty() && stat(path.c_str(), &st) == 0 && (double) st.st_mtime >= started - 120.0;
            }
            if (ours) {
                file_record = rec;
                // tell rank 0 that this rank has the record: it removes the marker once every rank has said so
                std::ofstream(output_path + kReadyMarker + ".ack" + std::to_string(config().world_rank)) << g_run_id << "\n";
                return true;
            }
        }
        usleep(2000);
    }
    return false;
}

static int compress(const std::string &folder_path, const std::string &output_path, double started) { // main.cpp:10-54
    const RunConfig &cfg = config();
    std::string file_record = (std::filesystem::path(folder_path).parent_path() / "sorted_files_by_size.txt").string();
    if (cfg.world_rank == 0) {
        std::cout << "Compressing folder: " << folder_path << std::endl;
        std::string tmp_record = sort_files_by_size(folder_path);
        std::cout << "File record saved location: " << tmp_record << std::endl;
        file_record = tmp_record;
        if (cfg.world_size > 1) publish_marker(output_path, file_record);
    } else {
        const char *to = std::getenv("ZWZ_RENDEZVOUS_TIMEOUT");
        if (!wait_for_marker(output_path, started, (to && *to) ? std::atof(to) : 600.0, file_record)) {
            std::cerr << "Rank: " << cfg.world_rank << " - timed out waiting for the file record of this run" << std::endl;
            return 5;
        }
    }
    std::cout << "file_record: " << file_record << std::endl;
    timing_mark("file record ready");
    int file_count = count_non_empty_lines(file_record);
    if (cfg.world_rank < file_count || cfg.world_size > 1) { // a rank without files of its own may still deflate segments of a big file
        do_compression(folder_path, output_path, file_record, cfg.world_rank);
    } else {
        std::cout << "Rank: " << cfg.world_rank << " - No file to compress" << std::endl;
    }
    std::cout << "main - Rank: " << cfg.world_rank << " - do_compression finished" << std::endl;
    if (cfg.world_rank == 0 && cfg.world_size > 1) {
        // every other rank acknowledges the marker when it has read it; only then may it go (a rank that is still looking for it
        // would otherwise wait for nothing). A rank that never shows up is a failed job: say so.
        const char *to = std::getenv("ZWZ_RENDEZVOUS_TIMEOUT");
        const double limit = (to && *to) ? std::atof(to) : 600.0, t0 = now_s();
        int missing = cfg.world_size - 1;
        while (missing > 0 && now_s() - t0 < limit) {
            missing = 0;
            for (int r = 1; r < cfg.world_size; ++r) {
                std::ifstream f(output_path + kReadyMarker + ".ack" + std::to_string(r));
                std::string id;
                if (!(f.is_open() && std::getline(f, id) && id == g_run_id)) ++missing;
            }
            if (missing) usleep(2000);
        }
        if (missing) {
            std::cerr << "Rank: 0 - " << missing << " rank(s) never picked up the file record of this run" << std::endl;
            return 5;
        }
        std::remove((output_path + kReadyMarker).c_str());
        for (int r = 1; r < cfg.world_size; ++r) std::remove((output_path + kReadyMarker + ".ack" + std::to_string(r)).c_str());
    }
    return 0;
}

// main.cpp:56-70 leaves decompression to rank 0 ("not supported in MPI parallel mode"). Here every rank takes its share of the
// batches (decompress_pipeline.cpp): `ZWZ_GPUS=8 main decompress <dir> <out>` uses the 8 GPUs of the box.
static void decompress(const std::string &source_path, const std::string &output_path) { do_decompression(source_path, output_path); }

namespace {

// rank 0 only (main.cpp:104-129): the source must exist; the output directory is created when missing
bool prepare_paths(const std::string &source_path, const std::string &output_path) {
    namespace fs = std::filesystem;
    std::error_code ec;
    if (!fs::exists(source_path, ec)) {
        std::cerr << "Source path does not exist.\n";
        return false;
    }
    if (!fs::exists(output_path, ec)) {
        if (mkdir(output_path.c_str(), 0777) == -1 && errno != EEXIST) { // EEXIST: another rank of this run was faster
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
    const bool decompressing = operation == "decompress";
    if (const char *id = std::getenv("ZWZ_RUN_ID")) g_run_id = id;
    cfg.run_id = g_run_id;
    if (cfg.world_rank == 0) {
        if (!prepare_paths(source_path, output_path)) return 1;
        if (compressing) { // whatever an earlier run left
            std::remove((output_path + kReadyMarker).c_str());
            for (int r = 1; r < std::max(cfg.world_size, 64); ++r) std::remove((output_path + kReadyMarker + ".ack" + std::to_string(r)).c_str());
        }
        if (cfg.world_size == 1 || g_run_id.empty()) { // the segment ledger of an earlier run (a launcher with ZWZ_RUN_ID may
            std::error_code ec;                        // already have ranks at work in it)
            std::filesystem::remove_all(output_path + "/.zwz_segments", ec);
        }
    }
    const char *g = std::getenv("ZWZ_GPUS");
    const int self_gpus = (g && cfg.world_size == 1) ? std::atoi(g) : 0;
    std::vector<pid_t> kids;
    if (self_gpus > 1 && (compressing || decompressing)) {
        if (g_run_id.empty()) g_run_id = "self-" + std::to_string((long) getpid()) + "-" + std::to_string((long long) (start_time * 1e6));
        cfg.run_id = g_run_id;
        kids = fork_ranks(self_gpus);
    }

    // the CUDA runtime and this rank's first context come up (0.6–2.5 s on a fresh 1-GPU box, ~10 s per rank on the 8-GPU box —
    // whether or not the other seven devices are hidden from the process: measured, profiles/round2/multi/n8_bringup_*.log) while
    // the host walks and sorts the tree or maps and indexes the archives
    std::thread warm;
    if (compressing || decompressing) warm = std::thread([] { zwzhost::warm_device(); });

    int rc = 0;
    try {
        if (compressing) {
            rc = compress(source_path, output_path, (double) std::time(nullptr));
        } else if (decompressing) {
            decompress(source_path, output_path);
        } else {
            std::cerr << "Invalid operation: " << operation << ". Please use 'compress' or 'decompress'.\n";
            return 1;
        }
    } catch (const std::exception &e) {
        std::cerr << "Rank: " << cfg.world_rank << " - fatal: " << e.what() << std::endl;
        rc = 2;
    }
    if (warm.joinable()) warm.join();
    timing_mark("operation finished");
    if (cfg.strict && stats().bad_records != 0 && rc == 0) {
        std::cerr << "ZWZ_STRICT: " << stats().bad_records << " record(s) did not decode cleanly" << std::endl;
        rc = 4;
    }
    if (self_gpus > 1 && cfg.world_rank != 0) { // a forked rank: no summary, no waiting
        std::cout << std::flush;
        std::cerr << std::flush;
        _exit(rc);
    }
    for (pid_t k : kids) {
        int st = 0;
        waitpid(k, &st, 0);
        if ((!WIFEXITED(st) || WEXITSTATUS(st) != 0) && rc == 0) rc = 3;
    }
    if (cfg.world_rank == 0) {
        if (!kids.empty()) { // every rank of a self-launched run has finished: the run's scratch files can go
            std::error_code ec;
            std::filesystem::remove_all(output_path + "/.zwz_segments", ec);
        }
        print_summary(operation, now_s() - start_time);
    }
    return rc;
}
