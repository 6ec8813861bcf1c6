// decompress_pipeline.cpp — the decompress side (decompression.cpp:45-178 in the reference).
//
// The reference parses an archive record by record and inflates each one on the spot (one thread per archive, serial
// inside a file), holding early records in a per-path min-heap. Here every archive is mapped and parsed into a record index
// first (`.zwz` has no index of its own: records are self-delimiting, decompression.cpp:65-92; only the header pages are
// touched), the reference's ordering rules are applied to that index, and whole files' worth of records form groups. W
// workers (pipeline.hpp) each take a group: copy its payloads from the mapping into page-locked staging, make ONE trip
// through the GPU (zwz_decompress_records: one upload, batched inflate, device-side concatenation per file, MD5 of every
// output file from the same resident bytes, one download) and write the group's files. Console lines are emitted in group
// order whatever W is.
//
// Reader rules kept from the reference:
//   * a file receives the records with sequence ids 0, 1, 2, ... up to the first missing id (later ones would wait in
//     the heap forever, decompression.cpp:119-153, "Warning: pending chunks remaining"); of records sharing an id the
//     first one in archive order is used;
//   * a path gets its output file as soon as it is seen (decompression.cpp:95-110), even if nothing is ever written;
//   * the MD5 verdict line is printed only when the record carrying is_last_chunk arrived in order and nothing was
//     pending behind it (decompression.cpp:132) — ZWZ_VERIFY_ALL=1 prints a verdict for every complete file instead;
//   * errors inside a stream are not fatal: whatever inflate produced is written (decompression.cpp:31).
#include "pipeline.hpp"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <iostream>
#include <condition_variable>
#include <map>
#include <memory>
#include <unordered_map>
#include <sstream>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace zwzhost {

namespace fs = std::filesystem;

namespace {

struct Rec {
    uint64_t off;   // payload offset inside the archive
    uint32_t len;   // payload length
    int seq;
    bool last;
};
struct FileState {
    std::string relpath;
    std::vector<Rec> recs;       // archive order
    std::string stored_md5;      // from the LAST record that carried one (file_md5s[path] is overwritten, decompression.cpp:91)
    bool verdict_in_order = false;
    std::vector<Rec> ordered;    // seq 0..k-1
    bool complete = false;       // the ordered run ends with an is_last record
};

// read-only view of a mapped archive
struct Span {
    const uint8_t *p = nullptr;
    size_t n = 0;
    const uint8_t *data() const { return p; }
    size_t size() const { return n; }
    uint8_t operator[](size_t i) const { return p[i]; }
};
struct Archive {
    std::string filename;
    Span bytes;
    void *map = nullptr;
    std::vector<FileState> files;
};
struct Group {
    size_t archive = 0, first = 0, count = 0; // files [first, first + count) of that archive
    size_t nrec = 0;
    bool segment = false;  // records [r0, r1) of ONE file whose records exceed a batch
    size_t r0 = 0, r1 = 0;
    int big = -1, seg = 0; // index into the big-file table, segment number inside the file
};

// decompression.cpp:65-92
void parse_archive(const Span &a, std::vector<FileState> &files) {
    struct Heap {                 // the reference's per-path reader state, simulated to know which verdicts it would print
        int expected = 0;         // expected_sequence_id
        std::vector<int> pending; // sequence ids sitting in its min-heap
    };
    std::unordered_map<std::string, size_t> index;
    std::vector<Heap> heaps;
    size_t cur = (size_t) -1; // records of one path are consecutive in every archive a writer produces: skip the lookup then
    size_t o = 0;
    const size_t n = a.size();
    while (o + 4 <= n) {
        int total_size, path_length;
        std::memcpy(&total_size, a.data() + o, 4);
        if (o + 8 > n) break;
        std::memcpy(&path_length, a.data() + o + 4, 4);
        if (path_length < 0 || o + 8 + (size_t) path_length + 5 > n) break;
        const char *path_bytes = (const char *) a.data() + o + 8;
        int sequence_id;
        std::memcpy(&sequence_id, a.data() + o + 8 + path_length, 4);
        bool is_last = a[o + 12 + path_length] != 0;
        long payload = (long) total_size - (4 + path_length + 4 + 1);
        size_t p0 = o + 13 + (size_t) path_length;
        if (payload < 0 || p0 + (size_t) payload > n) break;
        o = p0 + (size_t) payload;
        const char *md5_bytes = nullptr;
        size_t md5_len = 0;
        if (is_last) {
            md5_bytes = (const char *) a.data() + o;
            md5_len = o + MD5_DATA_SIZE > n ? n - o : MD5_DATA_SIZE;
            o += md5_len;
        }
        if (cur == (size_t) -1 || files[cur].relpath.size() != (size_t) path_length ||
            std::memcmp(files[cur].relpath.data(), path_bytes, (size_t) path_length) != 0) {
            std::string relpath(path_bytes, (size_t) path_length);
            auto it = index.find(relpath);
            if (it == index.end()) {
                it = index.emplace(relpath, files.size()).first;
                files.emplace_back();
                files.back().relpath = std::move(relpath);
                heaps.emplace_back();
            }
            cur = it->second;
        }
        FileState &fsx = files[cur];
        fsx.recs.push_back({(uint64_t) p0, (uint32_t) payload, sequence_id, is_last});
        if (is_last) fsx.stored_md5.assign(md5_bytes, md5_len);
        Heap &h = heaps[cur];
        if (h.expected == sequence_id) {
            ++h.expected;
            for (;;) {
                auto m = std::min_element(h.pending.begin(), h.pending.end());
                if (m == h.pending.end() || *m != h.expected) break;
                h.pending.erase(m);
                ++h.expected;
            }
            if (is_last && h.expected == sequence_id + 1 && h.pending.empty()) fsx.verdict_in_order = true;
        } else {
            h.pending.push_back(sequence_id);
        }
    }
    // the records a file receives: sequence ids 0, 1, 2, ... up to the first missing one, first occurrence of each id
    for (auto &f : files) {
        bool in_order = true;
        for (size_t i = 0; i < f.recs.size() && in_order; ++i) in_order = f.recs[i].seq == (int) i;
        if (in_order) { // what every writer produces; also keeps a 262 149-record file linear
            f.ordered = f.recs;
        } else {
            std::unordered_map<int, size_t> first;
            for (size_t i = 0; i < f.recs.size(); ++i) first.emplace(f.recs[i].seq, i);
            for (int want = 0;; ++want) {
                auto it = first.find(want);
                if (it == first.end()) break;
                f.ordered.push_back(f.recs[it->second]);
            }
        }
        f.complete = !f.ordered.empty() && f.ordered.back().last;
    }
}


// console text of one group
struct Console {
    std::ostringstream out, err;
};

// Console lines leave the process in group order whatever the workers' timing: a finished group posts its text, and whoever
// posts prints every consecutive group that is ready. Nobody ever waits here.
class OrderedConsole {
  public:
    void post(size_t g, std::string out, std::string err) {
        std::lock_guard<std::mutex> lock(mu_);
        ready_[g] = {std::move(out), std::move(err)};
        for (auto it = ready_.find(next_); it != ready_.end(); it = ready_.find(next_)) {
            if (!it->second.first.empty()) std::cout << it->second.first << std::flush;
            if (!it->second.second.empty()) std::cerr << it->second.second << std::flush;
            ready_.erase(it);
            ++next_;
        }
    }
    void skip(size_t g) { post(g, "", ""); } // a group that belongs to another rank

  private:
    std::mutex mu_;
    std::map<size_t, std::pair<std::string, std::string>> ready_;
    size_t next_ = 0;
};

// one directory-existence check per directory instead of one per file (files of a directory are consecutive in an archive)
void ensure_parent(const std::string &file_path, std::string &last_dir) {
    const size_t slash = file_path.find_last_of('/');
    if (slash == std::string::npos || slash == 0) return;
    if (last_dir.size() == slash && file_path.compare(0, slash, last_dir) == 0) return;
    last_dir.assign(file_path, 0, slash);
    std::error_code ec;
    if (!fs::exists(last_dir, ec)) fs::create_directories(last_dir, ec);
}

// decompression.cpp:140-146, plus the quarantine the reference's README promises (README.md:175,186) behind ZWZ_QUARANTINE=1:
// a file whose MD5 does not match is moved to <output dir>/bad/<relative path>
void print_verdict(Console &con, RunStats &st, const std::string &output_dir, const std::string &relpath, const std::string &stored,
                   const std::string &calculated) {
    const std::string file_path = output_dir + "/" + relpath;
    if (calculated != stored) {
        con.err << "MD5 mismatch for file: " << file_path << "\n";
        con.out << "Expected MD5: " << stored << "\n";
        con.out << "Calculated MD5: " << calculated << "\n";
        st.md5_mismatch++;
        if (config().quarantine) {
            const std::string bad = output_dir + "/bad/" + relpath;
            std::error_code ec;
            fs::create_directories(fs::path(bad).parent_path(), ec);
            fs::rename(file_path, bad, ec);
            if (!ec) con.err << "Moved to: " << bad << "\n";
        }
    } else {
        con.out << "MD5 match for file: " << file_path << "\n";
        st.md5_match++;
    }
}

const char *status_name(uint32_t s) {
    return s == ZWZ_STREAM_TRUNCATED ? "truncated stream" : (s == ZWZ_STREAM_BAD ? "invalid stream or checksum" : (s == ZWZ_STREAM_OUTPUT_FULL ? "output larger than expected" : "ok"));
}

std::mutex &stats_mu() {
    static std::mutex mu;
    return mu;
}
void merge_stats(const RunStats &s) {
    std::lock_guard<std::mutex> lock(stats_mu());
    RunStats &g = stats();
    g.files += s.files;
    g.records += s.records;
    g.raw_bytes += s.raw_bytes;
    g.md5_match += s.md5_match;
    g.md5_mismatch += s.md5_mismatch;
    g.bad_records += s.bad_records;
    g.t_read += s.t_read;
    g.t_gpu += s.t_gpu;
    g.t_write += s.t_write;
}

void write_all_at(int fd, const uint8_t *p, size_t n, uint64_t off) {
    while (n) {
        ssize_t w = ::pwrite(fd, p, n, (off_t) off);
        if (w < 0) {
            if (errno == EINTR) continue;
            throw std::runtime_error("zwz: short write to an output file");
        }
        p += w;
        n -= (size_t) w;
        off += (uint64_t) w;
    }
}

// A file whose records exceed a batch is cut into SEGMENTS of consecutive records that workers — of this rank, or of several
// ranks — inflate independently. A segment's place in the output file is the sum of the raw sizes of the segments before it,
// known only once those are inflated: the ledger hands out offsets in segment order. Inside one process it is a counter
// behind a condition variable; across ranks (ZWZ_GPUS=N decompress) the raw size of every finished segment is published as a
// tiny file under <output dir>/.zwz_segments/ and read by the ranks that need it (the same shared-directory channel the
// compress side uses for the file record; the reference format itself has no index to exchange).
struct BigOut {
    size_t archive = 0, file = 0;
    int nseg = 0;
    std::string path;   // output file
    std::string key;    // ledger key across ranks
    std::mutex mu;
    std::condition_variable cv;
    int next_seg = 0;          // single process: segments [0, next_seg) have their sizes in
    uint64_t next_off = 0;
    std::atomic<int> written{0};
    std::atomic<uint64_t> bad{0};
};

struct Job {
    std::vector<Archive> &archives;
    const std::vector<Group> &groups;
    std::vector<std::unique_ptr<BigOut>> &bigs;
    const std::string &output_dir;
    uint64_t budget;
    int device;
    int rank, world;
    size_t max_comp = 0, max_out = 0; // of any planned group: page-locked buffers are sized once per worker
    std::atomic<size_t> next_group{0};
    OrderedConsole console;
    OrderedCommit abort_only; // run_workers wants one; nothing waits on it here
    Job(std::vector<Archive> &a, const std::vector<Group> &g, std::vector<std::unique_ptr<BigOut>> &b, const std::string &out, uint64_t budget_,
        int device_, int rank_, int world_)
        : archives(a), groups(g), bigs(b), output_dir(out), budget(budget_), device(device_), rank(rank_), world(world_) {
        for (const auto &x : g) {
            size_t comp = 0;
            if (x.segment) {
                const auto &recs = a[x.archive].files[x.first].ordered;
                for (size_t r = x.r0; r < x.r1; ++r) comp += recs[r].len;
            } else {
                for (size_t f = x.first; f < x.first + x.count; ++f)
                    for (const auto &r : a[x.archive].files[f].ordered) comp += r.len;
            }
            max_comp = std::max(max_comp, comp);
            max_out = std::max(max_out, x.nrec * CHUNK_SIZE);
        }
    }
    bool mine(size_t g) const { return world <= 1 || (int) (g % (size_t) world) == rank; }
};

class Worker {
  public:
    Worker(Job &job, int id) : job_(job), ctx_(worker_ctx(job.device, id)), in_(ctx_), out_(ctx_) {}

    void run() {
        for (;;) {
            Pending &p = pend_[turn_++ & 1];
            finalize(p); // before a group is taken: the verdicts of two groups ago
            size_t g;
            for (;;) { // groups are dealt over the ranks; every rank walks the whole list
                g = job_.next_group.fetch_add(1);
                if (g >= job_.groups.size() || job_.mine(g)) break;
                job_.console.skip(g);
            }
            if (g >= job_.groups.size()) break;
            const Group &grp = job_.groups[g];
            p.g = g;
            if (grp.segment)
                segment(grp, p);
            else
                small_files(grp, p);
        }
        finalize(pend_[turn_ & 1]);
        finalize(pend_[(turn_ + 1) & 1]);
    }

  private:
    struct OutFile {
        FileState *fsx;
        uint64_t bytes;
        bool created;
    };
    // a group whose files are on disk, waiting for its digests to say what the console gets
    struct Pending {
        bool active = false;
        uint64_t ticket = 0;
        size_t g = 0;
        Console con;
        RunStats st;
        std::vector<uint8_t> digest;
        std::vector<OutFile> files;
    };

    void finalize(Pending &p) {
        if (!p.active) return;
        p.active = false;
        double t0 = now_seconds();
        if (p.ticket && zwz_wait(ctx_, p.ticket) != ZWZ_OK) throw std::runtime_error(std::string("zwz_wait: ") + zwz_last_error(ctx_));
        p.st.t_gpu += now_seconds() - t0;
        const RunConfig &cfg = config();
        for (size_t i = 0; i < p.files.size(); ++i) {
            FileState &fsx = *p.files[i].fsx;
            if (!p.files[i].created) continue;
            if (fsx.recs.size() != fsx.ordered.size()) p.con.err << "Warning: pending chunks remaining for file: " << fsx.relpath << "\n";
            if (cfg.verbose && !fsx.stored_md5.empty()) p.con.out << "Read MD5: " << fsx.stored_md5 << "\n"; // decompression.cpp:90
            if (fsx.complete && (fsx.verdict_in_order || cfg.verify_all)) {
                char hex[32];
                zwz_md5_hex(&p.digest[i * 16], hex);
                print_verdict(p.con, p.st, job_.output_dir, fsx.relpath, fsx.stored_md5, std::string(hex, 32));
            }
        }
        job_.console.post(p.g, p.con.out.str(), p.con.err.str());
        merge_stats(p.st);
        p.con = Console();
        p.st = RunStats();
        p.files.clear();
        p.ticket = 0;
    }

    // Records whose stream did not reach a clean end (truncated, invalid, checksum mismatch). Counted always; named on stderr
    // only with ZWZ_STRICT=1 — the reference writes whatever inflate produced and says nothing (decompression.cpp:31).
    void report_bad_records(const Archive &a, size_t first_file, const std::vector<uint32_t> &rfile, const std::vector<uint32_t> &status,
                            size_t first_rec, Console &con, RunStats &st) {
        size_t k = first_rec;
        uint32_t cur = 0xffffffffu;
        for (size_t i = 0; i < status.size(); ++i) {
            if (rfile[i] != cur) {
                cur = rfile[i];
                k = first_rec;
            }
            if (status[i] != ZWZ_STREAM_END) {
                st.bad_records++;
                const FileState &f = a.files[first_file + cur];
                if (config().strict) con.err << "Corrupt record: " << f.relpath << " sequence " << f.ordered[k].seq << " (" << status_name(status[i]) << ")\n";
            }
            ++k;
        }
    }

    // inflate the staged records into out_, retrying with larger capacities for foreign records that inflate to more than
    // 65 535 bytes (the reference's loop handles any size, decompression.cpp:17-33). digest != NULL: deferred (ticket)
    void inflate_group(std::vector<uint64_t> &off, std::vector<uint32_t> &len, std::vector<uint32_t> &rfile, uint32_t nf, std::vector<uint64_t> &foff,
                       uint8_t *digest, uint64_t *ticket, std::vector<uint32_t> &status) {
        const size_t nrec = off.size();
        std::vector<uint32_t> cap(nrec, (uint32_t) CHUNK_SIZE), raw_len(nrec);
        status.assign(nrec, 0u);
        for (int attempt = 0;; ++attempt) {
            uint64_t need = 0;
            for (size_t i = 0; i < nrec; ++i) need += cap[i];
            out_.reserve(std::max<size_t>(need, job_.max_out) + 64);
            int rc = digest ? zwz_decompress_records_async(ctx_, in_.data(), off.data(), len.data(), cap.data(), rfile.data(), (uint32_t) nrec, nf,
                                                           out_.data(), out_.cap, foff.data(), raw_len.data(), status.data(), digest, 0, ticket)
                            : zwz_decompress_records(ctx_, in_.data(), off.data(), len.data(), cap.data(), rfile.data(), (uint32_t) nrec, nf, out_.data(),
                                                     out_.cap, foff.data(), raw_len.data(), status.data(), nullptr, 0);
            if (rc != ZWZ_OK) throw std::runtime_error(std::string("zwz_decompress_records: ") + zwz_last_error(ctx_));
            bool again = false;
            for (size_t i = 0; i < nrec; ++i)
                if (status[i] == ZWZ_STREAM_OUTPUT_FULL) {
                    cap[i] = raw_len[i];
                    again = true;
                }
            if (!again || attempt >= 2) break;
        }
    }

    // copies the payloads of `n` records (described by src[k] = archive offset, len[k]) into in_ back to back
    void stage_payloads(const Span &arch, const std::vector<uint64_t> &src, const std::vector<uint64_t> &dst, const std::vector<uint32_t> &len) {
        parallel_for(src.size(), io_threads(), [&](size_t k) { std::memcpy(in_.data() + dst[k], arch.data() + src[k], len[k]); });
    }

    void small_files(const Group &grp, Pending &p) {
        Archive &a = job_.archives[grp.archive];
        RunStats &st = p.st;
        const uint32_t nf = (uint32_t) grp.count;
        double t0 = now_seconds();
        std::vector<uint64_t> src, off, foff(nf + 1);
        std::vector<uint32_t> len, rfile;
        src.reserve(grp.nrec);
        off.reserve(grp.nrec);
        len.reserve(grp.nrec);
        rfile.reserve(grp.nrec);
        uint64_t used = 0;
        for (size_t f = grp.first; f < grp.first + grp.count; ++f)
            for (const auto &r : a.files[f].ordered) {
                src.push_back(r.off);
                off.push_back(used);
                len.push_back(r.len);
                rfile.push_back((uint32_t) (f - grp.first));
                used += r.len;
            }
        in_.reserve(std::max<size_t>(used, job_.max_comp) + 64);
        stage_payloads(a.bytes, src, off, len);
        st.t_read += now_seconds() - t0;
        p.digest.assign((size_t) nf * 16, 0);
        t0 = now_seconds();
        std::vector<uint32_t> status;
        if (!off.empty()) {
            inflate_group(off, len, rfile, nf, foff, p.digest.data(), &p.ticket, status);
            report_bad_records(a, grp.first, rfile, status, 0, p.con, st);
        } else if (nf) { // files without a single usable record: still created, and their verdict is that of the empty file
            zwz_md5_batch(ctx_, in_.data(), foff.data(), foff.data(), nf, p.digest.data());
        }
        st.t_gpu += now_seconds() - t0;
        // ---- write: directories first (serial, one check per directory), then the files on several threads
        t0 = now_seconds();
        for (size_t f = grp.first; f < grp.first + grp.count; ++f) ensure_parent(job_.output_dir + "/" + a.files[f].relpath, last_dir_);
        p.files.assign(nf, OutFile{nullptr, 0, false});
        parallel_for(nf, io_threads(), [&](size_t i) {
            FileState &fsx = a.files[grp.first + i];
            const std::string file_path = job_.output_dir + "/" + fsx.relpath;
            const uint64_t a0 = foff[i], a1 = foff[i + 1];
            p.files[i].fsx = &fsx;
            p.files[i].bytes = a1 - a0;
            int fd = ::open(file_path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
            if (fd < 0) return;
            try {
                if (a1 > a0) write_all_at(fd, out_.data() + a0, (size_t) (a1 - a0), 0);
                p.files[i].created = true;
            } catch (...) {
            }
            ::close(fd);
        });
        for (uint32_t i = 0; i < nf; ++i) {
            if (!p.files[i].created) {
                p.con.err << "Error creating output file: " << job_.output_dir << "/" << p.files[i].fsx->relpath << "\n";
                continue;
            }
            st.files++;
            st.records += p.files[i].fsx->ordered.size();
            st.raw_bytes += p.files[i].bytes;
        }
        st.t_write += now_seconds() - t0;
        p.active = true;
    }

    // One segment of a file that does not fit a batch: inflate (no digest), learn the offset from the ledger, write. Whoever
    // writes the last missing segment hashes the finished file the way the reference does — by reading it back
    // (decompression.cpp:136) — streamed through the GPU.
    void segment(const Group &grp, Pending &p) {
        Archive &a = job_.archives[grp.archive];
        FileState &fsx = a.files[grp.first];
        BigOut &big = *job_.bigs[(size_t) grp.big];
        RunStats &st = p.st;
        double t0 = now_seconds();
        std::vector<uint64_t> src, off, foff(2);
        std::vector<uint32_t> len, rfile(grp.r1 - grp.r0, 0u);
        uint64_t used = 0;
        for (size_t r = grp.r0; r < grp.r1; ++r) {
            src.push_back(fsx.ordered[r].off);
            off.push_back(used);
            len.push_back(fsx.ordered[r].len);
            used += fsx.ordered[r].len;
        }
        in_.reserve(std::max<size_t>(used, job_.max_comp) + 64);
        stage_payloads(a.bytes, src, off, len);
        st.t_read += now_seconds() - t0;
        t0 = now_seconds();
        std::vector<uint32_t> status;
        inflate_group(off, len, rfile, 1, foff, nullptr, nullptr, status);
        report_bad_records(a, grp.first, rfile, status, grp.r0, p.con, st);
        st.t_gpu += now_seconds() - t0;
        const uint64_t bytes = foff[1];
        // ---- the segment's place in the file
        t0 = now_seconds();
        uint64_t at = 0;
        if (job_.world > 1) {
            ledger_publish(job_.output_dir, big.key, grp.seg, ".size", bytes);
            for (int s = 0; s < grp.seg; ++s) at += ledger_wait(job_.output_dir, big.key, s, ".size");
        } else {
            std::unique_lock<std::mutex> lock(big.mu);
            big.cv.wait(lock, [&] { return big.next_seg == grp.seg; });
            at = big.next_off;
            big.next_off += bytes;
            big.next_seg = grp.seg + 1;
            lock.unlock();
            big.cv.notify_all();
        }
        int fd = ::open(big.path.c_str(), O_WRONLY | O_CREAT, 0666);
        bool ok = fd >= 0;
        if (ok) {
            if (bytes) write_all_at(fd, out_.data(), (size_t) bytes, at);
        } else if (grp.seg == 0) {
            p.con.err << "Error creating output file: " << big.path << "\n";
        }
        st.records += grp.r1 - grp.r0;
        st.raw_bytes += bytes;
        st.t_write += now_seconds() - t0;
        // ---- the last segment to land finishes the file
        bool finisher;
        if (job_.world > 1) {
            ledger_publish(job_.output_dir, big.key, grp.seg, ".done", bytes);
            finisher = grp.seg == big.nseg - 1;
            if (finisher)
                for (int s = 0; s < big.nseg; ++s) ledger_wait(job_.output_dir, big.key, s, ".done");
        } else {
            finisher = big.written.fetch_add(1) + 1 == big.nseg;
        }
        if (finisher && ok) {
            const uint64_t total = job_.world > 1 ? at + bytes : big.next_off;
            if (ftruncate(fd, (off_t) total) != 0) p.con.err << "Error truncating output file: " << big.path << "\n"; // an older, longer file of that name
        }
        if (fd >= 0) ::close(fd);
        if (finisher) {
            st.files++;
            if (fsx.recs.size() != fsx.ordered.size()) p.con.err << "Warning: pending chunks remaining for file: " << fsx.relpath << "\n";
            if (config().verbose && !fsx.stored_md5.empty()) p.con.out << "Read MD5: " << fsx.stored_md5 << "\n";
            if (ok && fsx.complete && (fsx.verdict_in_order || config().verify_all)) {
                t0 = now_seconds();
                print_verdict(p.con, st, job_.output_dir, fsx.relpath, fsx.stored_md5, md5_of_file_ctx(ctx_, big.path));
                st.t_gpu += now_seconds() - t0;
            }
        }
        p.active = true; // nothing deferred, but the console text goes out through the same door
    }

    Job &job_;
    zwz_ctx *ctx_;
    PinnedBuf in_, out_;
    std::string last_dir_;
    Pending pend_[2];
    unsigned turn_ = 0;
};

bool map_archive(Archive &a) {
    int fd = ::open(a.filename.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat sb {};
    if (fstat(fd, &sb) != 0) {
        ::close(fd);
        return false;
    }
    a.bytes.n = (size_t) sb.st_size;
    if (a.bytes.n) {
        a.map = mmap(nullptr, a.bytes.n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (a.map == MAP_FAILED) {
            a.map = nullptr;
            ::close(fd);
            return false;
        }
        a.bytes.p = (const uint8_t *) a.map;
    }
    ::close(fd);
    return true;
}

} // namespace

// process.hpp:40. The reference runs one thread per archive (decompression.cpp:174) on rank 0 only (main.cpp:61-69); here the
// groups of all archives feed one worker pool per rank, and with several ranks (ZWZ_GPUS=N or an external launcher) the
// groups — whole small files, or record ranges of a large one — are dealt round-robin over the ranks' GPUs.
void do_decompression(const std::string &input_dir, const std::string &output_dir) {
    const RunConfig &cfg = config();
    std::vector<std::string> names;
    for (const auto &entry : fs::directory_iterator(input_dir)) {
        if (entry.path().extension() == ".zwz") names.push_back(entry.path().string()); // decompression.cpp:168-172
    }
    std::sort(names.begin(), names.end());

    // the device comes up (CUDA runtime + first context: 0.5–2.5 s depending on how many GPUs are visible) while the archives
    // are mapped and indexed
    std::exception_ptr warm_error;
    std::thread warm([&] {
        try {
            if (!names.empty()) ctx_for(cfg.device);
        } catch (...) {
            warm_error = std::current_exception();
        }
    });
    double t0 = now_seconds();
    std::vector<Archive> archives(names.size());
    std::vector<char> opened(names.size(), 0);
    parallel_for(names.size(), io_threads() * 2, [&](size_t i) { // archives are independent: index them side by side
        archives[i].filename = names[i];
        if (!map_archive(archives[i])) return;
        opened[i] = 1;
        parse_archive(archives[i].bytes, archives[i].files);
    });
    {
        std::vector<Archive> kept;
        for (size_t i = 0; i < archives.size(); ++i) {
            if (!opened[i]) {
                std::cerr << "Error opening file: " << names[i] << std::endl;
                continue;
            }
            kept.push_back(std::move(archives[i]));
        }
        archives.swap(kept);
    }
    // groups of whole files, bounded by the raw bytes they may produce; batches small enough that every worker (of every rank)
    // gets several
    uint64_t total_rec = 0;
    for (const auto &a : archives)
        for (const auto &f : a.files) total_rec += f.ordered.size();
    const int pool = worker_count();
    const uint64_t even = total_rec * CHUNK_SIZE / ((uint64_t) pool * 3 * (uint64_t) std::max(1, cfg.world_size)) + 1;
    const uint64_t budget = std::max<uint64_t>(std::min<uint64_t>(cfg.batch_bytes, std::max<uint64_t>(even, (uint64_t) 8 << 20)), 4 * CHUNK_SIZE);
    const size_t per_seg = (size_t) std::max<uint64_t>(1, budget / CHUNK_SIZE);
    std::vector<Group> groups;
    std::vector<std::unique_ptr<BigOut>> bigs;
    for (size_t ai = 0; ai < archives.size(); ++ai) {
        const auto &files = archives[ai].files;
        size_t fi = 0;
        while (fi < files.size()) {
            if (files[fi].ordered.size() * CHUNK_SIZE > budget) {
                auto big = std::make_unique<BigOut>();
                big->archive = ai;
                big->file = fi;
                big->path = output_dir + "/" + files[fi].relpath;
                big->key = "a" + std::to_string(ai) + "f" + std::to_string(fi) + "_" + cfg.run_id;
                const size_t nrec = files[fi].ordered.size();
                for (size_t r0 = 0; r0 < nrec; r0 += per_seg) {
                    Group g;
                    g.archive = ai;
                    g.first = fi;
                    g.count = 1;
                    g.segment = true;
                    g.r0 = r0;
                    g.r1 = std::min(nrec, r0 + per_seg);
                    g.nrec = g.r1 - g.r0;
                    g.big = (int) bigs.size();
                    g.seg = big->nseg++;
                    groups.push_back(g);
                }
                bigs.push_back(std::move(big));
                ++fi;
                continue;
            }
            size_t fj = fi, nrec = 0;
            uint64_t est = 0;
            while (fj < files.size() && (fj == fi || est + files[fj].ordered.size() * CHUNK_SIZE <= budget) &&
                   files[fj].ordered.size() * CHUNK_SIZE <= budget) {
                est += files[fj].ordered.size() * CHUNK_SIZE;
                nrec += files[fj].ordered.size();
                ++fj;
            }
            Group g;
            g.archive = ai;
            g.first = fi;
            g.count = fj - fi;
            g.nrec = nrec;
            groups.push_back(g);
            fi = fj;
        }
    }
    stats().t_read += now_seconds() - t0;
    if (!bigs.empty()) {
        std::string last;
        for (auto &b : bigs) ensure_parent(b->path, last);
        if (cfg.world_size > 1) {
            std::error_code ec;
            fs::create_directories(ledger_dir(output_dir), ec);
        }
    }
    warm.join();
    if (warm_error) std::rethrow_exception(warm_error);

    Job job(archives, groups, bigs, output_dir, budget, cfg.device, cfg.world_rank, cfg.world_size);
    size_t mine = 0;
    for (size_t g = 0; g < groups.size(); ++g) mine += job.mine(g);
    const int workers = (int) std::min<size_t>((size_t) pool, std::max<size_t>(1, mine));
    if (mine)
        run_workers(workers, job.abort_only, [&](int w) {
            Worker worker(job, w);
            worker.run();
        });
    for (auto &a : archives)
        if (a.map) munmap(a.map, a.bytes.n);
    print_timing("decompress");
}

} // namespace zwzhost
