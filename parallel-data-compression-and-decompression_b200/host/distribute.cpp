// distribute.cpp — work distribution: walk, size-descending sort, record file (file_sort.cpp:14-43, file_tools.cpp:6-23).
//
// The deal itself (rank r takes sorted files r, r+P, ...; compression.cpp:31-41) is the multi-GPU sharding rule of this
// build: GPU index in place of MPI rank, no payload ever crosses GPUs.
#include "zwz_host.hpp"
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>

namespace zwzhost {

namespace fs = std::filesystem;

RunConfig &config() {
    static RunConfig c;
    return c;
}
RunStats &stats() {
    static RunStats s;
    return s;
}

double now_seconds() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}
static double g_t0 = now_seconds();
int timing_level() {
    const char *e = std::getenv("ZWZ_TIMING");
    return (e && *e) ? std::atoi(e) : 0;
}
void timing_mark(const char *what) { // ZWZ_TIMING>=1: seconds since the process started, at a named point
    if (timing_level() >= 1) std::cerr << "[zwz timing] t+" << now_seconds() - g_t0 << " s: " << what << std::endl;
}
zwz_ctx *ctx_for(int device);
void print_timing(const char *what) {
    if (timing_level() < 1) return;
    const RunStats &s = stats();
    std::cerr << "[zwz timing] " << what << ": init " << s.t_init << " s, read/parse " << s.t_read << " s, gpu passes " << s.t_gpu
              << " s, write " << s.t_write << " s; " << s.files << " files, " << s.records << " records, " << s.raw_bytes << " raw bytes"
              << std::endl;
    if (timing_level() >= 2) { // per-kernel device time (events around every launch; enabled in ctx_for)
        double ms[ZWZ_PROF_N];
        uint64_t launches[ZWZ_PROF_N];
        if (zwz_profile_read(ctx_for(config().device), ms, launches, 0) == ZWZ_OK)
            std::cerr << "[zwz timing] kernels (ms/launches): match " << ms[ZWZ_PROF_MATCH] << "/" << launches[ZWZ_PROF_MATCH] << ", encode "
                      << ms[ZWZ_PROF_ENCODE] << "/" << launches[ZWZ_PROF_ENCODE] << ", inflate " << ms[ZWZ_PROF_INFLATE] << "/"
                      << launches[ZWZ_PROF_INFLATE] << ", md5 " << ms[ZWZ_PROF_MD5] << "/" << launches[ZWZ_PROF_MD5] << ", pack/gather "
                      << ms[ZWZ_PROF_PACK] << "/" << launches[ZWZ_PROF_PACK] << std::endl;
    }
}

// GPUs this process may use, without initialising CUDA: the entries of CUDA_VISIBLE_DEVICES if set, else the /dev/nvidia<N>
// nodes; the CUDA runtime is asked only if neither says anything.
int visible_gpu_count() {
    if (const char *v = std::getenv("CUDA_VISIBLE_DEVICES")) {
        if (!*v) return 0;
        int n = 1;
        for (const char *p = v; *p; ++p) n += *p == ',';
        return n;
    }
    int n = 0;
    std::error_code ec;
    for (const auto &e : fs::directory_iterator("/dev", ec)) {
        const std::string name = e.path().filename().string();
        if (name.size() > 6 && name.compare(0, 6, "nvidia") == 0 && std::all_of(name.begin() + 6, name.end(), [](unsigned char ch) { return std::isdigit(ch) != 0; }))
            ++n;
    }
    return n > 0 ? n : zwz_device_count();
}

static int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

// The reference learns rank/size from MPI (main.cpp:12-13). Without MPI the launcher's environment carries them:
// ZWZ_RANK/ZWZ_WORLD (ours) or the Open MPI / PMI variables, so `mpirun -n P main ...` keeps working wherever an MPI launcher
// exists. The generic RANK/WORLD_SIZE/LOCAL_RANK of torchrun-style launchers are honoured only with ZWZ_LAUNCHER=env: a
// plain `main compress` inside a job pod that happens to export them must not silently turn into "rank r of N" (the
// reference never shards without mpirun).
void config_from_env() {
    RunConfig &c = config();
    const char *l = std::getenv("ZWZ_LAUNCHER");
    const bool generic = l && std::string(l) == "env";
    c.world_rank = env_int("ZWZ_RANK", env_int("OMPI_COMM_WORLD_RANK", env_int("PMI_RANK", generic ? env_int("RANK", 0) : 0)));
    c.world_size = env_int("ZWZ_WORLD", env_int("OMPI_COMM_WORLD_SIZE", env_int("PMI_SIZE", generic ? env_int("WORLD_SIZE", 1) : 1)));
    int local = env_int("ZWZ_LOCAL_RANK", env_int("OMPI_COMM_WORLD_LOCAL_RANK", generic ? env_int("LOCAL_RANK", c.world_rank) : c.world_rank));
    if (c.world_size < 1) c.world_size = 1;
    if (c.world_rank < 0 || c.world_rank >= c.world_size) c.world_rank = 0;
    if (c.world_size > 1)
        std::cerr << "zwz: running as rank " << c.world_rank << " of " << c.world_size << " (from the launcher's environment)" << std::endl;
    // Local rank 0 needs no device count (device 0), and the count is taken without touching CUDA where possible: the runtime
    // then initialises on a helper thread while the host walks directories / indexes archives, and ZWZ_GPUS can fork before any
    // CUDA call has been made.
    int ndev = local > 0 ? visible_gpu_count() : 1;
    c.device = env_int("ZWZ_DEVICE", ndev > 0 ? local % ndev : 0);
    c.level = env_int("ZWZ_LEVEL", 0);
    c.verbose = env_int("ZWZ_VERBOSE", 0) != 0;
    c.verify_all = env_int("ZWZ_VERIFY_ALL", 0) != 0;
    c.strict = env_int("ZWZ_STRICT", 0) != 0;
    c.quarantine = env_int("ZWZ_QUARANTINE", 0) != 0;
    int mb = env_int("ZWZ_BATCH_MB", 0);
    if (mb > 0) c.batch_bytes = (std::size_t) mb << 20;
}

std::vector<FileEntry> collect_and_sort(const fs::path &path) {
    std::vector<FileEntry> files;
    // Same iterator as file_sort.cpp:15-20 (the iteration order feeds the unstable sort below, so it decides how ties are
    // dealt). The relative path is the tail of the entry's path behind the root — what fs::relative() returns for entries
    // found below `path`, without normalising each of 370 000 paths (2 s of a 2.7 s walk on the README-shaped tree).
    const std::string root = path.string();
    const size_t cut = root.size() + ((!root.empty() && root.back() == '/') ? 0 : 1);
    for (const auto &entry : fs::recursive_directory_iterator(path)) {
        if (!entry.is_regular_file()) continue;
        const std::string &full = entry.path().native();
        std::string rel = (full.size() > cut && full.compare(0, root.size(), root) == 0) ? full.substr(cut) : fs::relative(entry.path(), path).string();
        files.push_back({std::move(rel), static_cast<off_t>(entry.file_size())});
    }
    // same comparator, same (unstable) algorithm as file_sort.cpp:30-31: ties land where libstdc++'s introsort puts them,
    // so on one box the deal is identical to the reference's
    std::sort(files.begin(), files.end(), [](const FileEntry &a, const FileEntry &b) { return a.size > b.size; });
    return files;
}

std::string sort_files_by_size(const fs::path &path) {
    std::vector<FileEntry> files = collect_and_sort(path);
    auto output_filename = path.parent_path() / "sorted_files_by_size.txt"; // file_sort.cpp:33
    // written beside its place and renamed into it: a rank that opens the record file sees all of it or the previous one, never
    // half of it
    const std::string tmp = output_filename.string() + ".tmp" + std::to_string((long) getpid());
    {
        std::ofstream file(tmp);
        if (file.is_open()) {
            for (const auto &e : files) file << e.relpath << "\n";
        }
    }
    if (std::rename(tmp.c_str(), output_filename.c_str()) != 0) std::remove(tmp.c_str());
    return output_filename.string();
}

int count_non_empty_lines(const std::string &file_path) {
    std::ifstream file(file_path);
    if (!file.is_open()) {
        std::cerr << "Error opening file: " << file_path << std::endl;
        return -1;
    }
    std::string line;
    int lines = 0;
    while (std::getline(file, line)) {
        bool blank = true;
        for (unsigned char ch : line)
            if (!std::isspace(ch)) {
                blank = false;
                break;
            }
        if (!blank) ++lines;
    }
    return lines;
}

} // namespace zwzhost
