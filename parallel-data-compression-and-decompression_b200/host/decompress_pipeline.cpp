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
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <iostream>
#include <map>
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
    bool big = false; // one file whose records exceed a batch: streamed through in sub-batches
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


// `last_dir` remembers the directory of the previous call (files of one directory are consecutive in an archive): one
// existence check per directory instead of one per file
void ensure_parent(const std::string &file_path, std::string &last_dir) {
    const size_t slash = file_path.find_last_of('/');
    if (slash == std::string::npos || slash == 0) return;
    if (last_dir.size() == slash && file_path.compare(0, slash, last_dir) == 0) return;
    last_dir.assign(file_path, 0, slash);
    std::error_code ec;
    if (!fs::exists(last_dir, ec)) fs::create_directories(last_dir, ec);
}

// console text of one group, released in group order
struct Console {
    std::ostringstream out, err;
};

void print_verdict(Console &con, RunStats &st, const std::string &file_path, const std::string &stored, const std::string &calculated) {
    if (calculated != stored) { // decompression.cpp:140-146
        con.err << "MD5 mismatch for file: " << file_path << "\n";
        con.out << "Expected MD5: " << stored << "\n";
        con.out << "Calculated MD5: " << calculated << "\n";
        st.md5_mismatch++;
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

struct Job {
    std::vector<Archive> &archives;
    const std::vector<Group> &groups;
    const std::string &output_dir;
    uint64_t budget;
    int device;
    size_t max_comp = 0, max_out = 0; // of any planned group: page-locked buffers are sized once per worker
    std::atomic<size_t> next_group{0};
    OrderedCommit order;
    Job(std::vector<Archive> &a, const std::vector<Group> &g, const std::string &out, uint64_t budget_, int device_)
        : archives(a), groups(g), output_dir(out), budget(budget_), device(device_) {
        for (const auto &x : g) {
            if (x.big) continue;
            size_t comp = 0;
            for (size_t f = x.first; f < x.first + x.count; ++f)
                for (const auto &r : a[x.archive].files[f].ordered) comp += r.len;
            max_comp = std::max(max_comp, comp);
            max_out = std::max(max_out, x.nrec * CHUNK_SIZE);
        }
    }
};

class Worker {
  public:
    Worker(Job &job, int id) : job_(job), ctx_(worker_ctx(job.device, id)), in_(ctx_), out_(ctx_) {}

    void run() {
        for (;;) {
            size_t g = job_.next_group.fetch_add(1);
            if (g >= job_.groups.size()) return;
            const Group &grp = job_.groups[g];
            Console con;
            RunStats st;
            if (grp.big)
                big_file(job_.archives[grp.archive], job_.archives[grp.archive].files[grp.first], con, st);
            else
                small_files(grp, con, st);
            job_.order.wait_turn(g);
            std::string o = con.out.str(), e = con.err.str();
            if (!o.empty()) std::cout << o << std::flush;
            if (!e.empty()) std::cerr << e << std::flush;
            job_.order.done(g);
            merge_stats(st);
        }
    }

  private:
    // Records whose stream did not reach a clean end (truncated, invalid, checksum mismatch). Counted always; named on stderr
    // only with ZWZ_STRICT=1 — the reference writes whatever inflate produced and says nothing (decompression.cpp:31).
    void report_bad_records(const Archive &a, size_t first_file, const std::vector<uint32_t> &rfile, const std::vector<uint32_t> &status,
                            Console &con, RunStats &st) {
        size_t k = 0;
        uint32_t cur = 0xffffffffu;
        for (size_t i = 0; i < status.size(); ++i) {
            if (rfile[i] != cur) {
                cur = rfile[i];
                k = 0;
            }
            if (status[i] != ZWZ_STREAM_END) {
                st.bad_records++;
                const FileState &f = a.files[first_file + cur];
                if (config().strict) con.err << "Corrupt record: " << f.relpath << " sequence " << f.ordered[k].seq << " (" << status_name(status[i]) << ")\n";
            }
            ++k;
        }
    }

    // inflate `nrec` records (payloads already compacted in in_) into out_, retrying with larger capacities for foreign
    // records that inflate to more than 65 535 bytes (the reference's loop handles any size, decompression.cpp:17-33)
    void inflate_group(std::vector<uint64_t> &off, std::vector<uint32_t> &len, std::vector<uint32_t> &rfile, uint32_t nf,
                       std::vector<uint64_t> &foff, uint8_t *digest, std::vector<uint32_t> &status) {
        const size_t nrec = off.size();
        std::vector<uint32_t> cap(nrec, (uint32_t) CHUNK_SIZE), raw_len(nrec);
        status.assign(nrec, 0u);
        for (int attempt = 0;; ++attempt) {
            uint64_t need = 0;
            for (size_t i = 0; i < nrec; ++i) need += cap[i];
            out_.reserve(need + 64);
            int rc = zwz_decompress_records(ctx_, in_.data(), off.data(), len.data(), cap.data(), rfile.data(), (uint32_t) nrec, nf, out_.data(),
                                            out_.cap, foff.data(), raw_len.data(), status.data(), digest, 0);
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

    // copies the payloads of records [r0, r1) of `recs` into in_ back to back, appending their offsets/lengths
    void stage_payloads(const Span &arch, const std::vector<Rec> &recs, size_t r0, size_t r1, uint64_t &used, std::vector<uint64_t> &off,
                        std::vector<uint32_t> &len) {
        for (size_t r = r0; r < r1; ++r) {
            std::memcpy(in_.data() + used, arch.data() + recs[r].off, recs[r].len);
            off.push_back(used);
            len.push_back(recs[r].len);
            used += recs[r].len;
        }
    }

    void small_files(const Group &grp, Console &con, RunStats &st) {
        Archive &a = job_.archives[grp.archive];
        const uint32_t nf = (uint32_t) grp.count;
        double t0 = now_seconds();
        uint64_t comp_bytes = 0;
        for (size_t f = grp.first; f < grp.first + grp.count; ++f)
            for (const auto &r : a.files[f].ordered) comp_bytes += r.len;
        in_.reserve(std::max<size_t>(comp_bytes, job_.max_comp) + 64);
        out_.reserve(job_.max_out + 64);
        std::vector<uint64_t> off, foff(nf + 1);
        std::vector<uint32_t> len, rfile;
        off.reserve(grp.nrec);
        len.reserve(grp.nrec);
        rfile.reserve(grp.nrec);
        uint64_t used = 0;
        for (size_t f = grp.first; f < grp.first + grp.count; ++f) {
            stage_payloads(a.bytes, a.files[f].ordered, 0, a.files[f].ordered.size(), used, off, len);
            rfile.insert(rfile.end(), a.files[f].ordered.size(), (uint32_t) (f - grp.first));
        }
        st.t_read += now_seconds() - t0;
        std::vector<uint8_t> digest((size_t) nf * 16);
        t0 = now_seconds();
        std::vector<uint32_t> status;
        if (!off.empty()) {
            inflate_group(off, len, rfile, nf, foff, digest.data(), status);
            report_bad_records(a, grp.first, rfile, status, con, st);
        }
        else if (nf) // files without a single usable record: still created, and their verdict is that of the empty file
            zwz_md5_batch(ctx_, in_.data(), foff.data(), foff.data(), nf, digest.data());
        st.t_gpu += now_seconds() - t0;
        t0 = now_seconds();
        const RunConfig &cfg = config();
        for (size_t f = grp.first; f < grp.first + grp.count; ++f) {
            FileState &fsx = a.files[f];
            std::string file_path = job_.output_dir + "/" + fsx.relpath;
            ensure_parent(file_path, last_dir_);
            std::FILE *o = std::fopen(file_path.c_str(), "wb");
            if (!o) {
                con.err << "Error creating output file: " << file_path << "\n";
                continue;
            }
            uint64_t a0 = foff[f - grp.first], a1 = foff[f - grp.first + 1];
            if (a1 > a0) std::fwrite(out_.data() + a0, 1, (size_t) (a1 - a0), o);
            std::fclose(o);
            st.files++;
            st.records += fsx.ordered.size();
            st.raw_bytes += a1 - a0;
            if (fsx.recs.size() != fsx.ordered.size()) con.err << "Warning: pending chunks remaining for file: " << fsx.relpath << "\n";
            if (cfg.verbose && !fsx.stored_md5.empty()) con.out << "Read MD5: " << fsx.stored_md5 << "\n"; // decompression.cpp:90
            if (fsx.complete && (fsx.verdict_in_order || cfg.verify_all)) {
                char hex[32];
                zwz_md5_hex(&digest[(f - grp.first) * 16], hex);
                print_verdict(con, st, file_path, fsx.stored_md5, std::string(hex, 32));
            }
        }
        st.t_write += now_seconds() - t0;
    }

    // A file whose records do not fit one batch: sub-batches of consecutive records are appended to the output file and the
    // MD5 is taken the way the reference takes it — by reading the finished file back (decompression.cpp:136).
    void big_file(Archive &a, FileState &fsx, Console &con, RunStats &st) {
        const RunConfig &cfg = config();
        std::string file_path = job_.output_dir + "/" + fsx.relpath;
        ensure_parent(file_path, last_dir_);
        std::FILE *o = std::fopen(file_path.c_str(), "wb");
        if (!o) {
            con.err << "Error creating output file: " << file_path << "\n";
            return;
        }
        const size_t per = std::max<size_t>(1, job_.budget / CHUNK_SIZE);
        uint64_t written = 0;
        for (size_t r0 = 0; r0 < fsx.ordered.size(); r0 += per) {
            size_t r1 = std::min(fsx.ordered.size(), r0 + per);
            uint64_t comp_bytes = 0;
            for (size_t r = r0; r < r1; ++r) comp_bytes += fsx.ordered[r].len;
            in_.reserve(comp_bytes + 64);
            std::vector<uint64_t> off, foff(2);
            std::vector<uint32_t> len, rfile(r1 - r0, 0u);
            uint64_t used = 0;
            stage_payloads(a.bytes, fsx.ordered, r0, r1, used, off, len);
            std::vector<uint32_t> status;
            inflate_group(off, len, rfile, 1, foff, nullptr, status);
            for (size_t i = 0; i < status.size(); ++i)
                if (status[i] != ZWZ_STREAM_END) {
                    st.bad_records++;
                    if (config().strict) con.err << "Corrupt record: " << fsx.relpath << " sequence " << fsx.ordered[r0 + i].seq << " (" << status_name(status[i]) << ")\n";
                }
            if (foff[1]) std::fwrite(out_.data(), 1, (size_t) foff[1], o);
            written += foff[1];
        }
        std::fclose(o);
        st.files++;
        st.records += fsx.ordered.size();
        st.raw_bytes += written;
        if (fsx.recs.size() != fsx.ordered.size()) con.err << "Warning: pending chunks remaining for file: " << fsx.relpath << "\n";
        if (fsx.complete && (fsx.verdict_in_order || cfg.verify_all))
            print_verdict(con, st, file_path, fsx.stored_md5, md5_of_file_ctx(ctx_, file_path));
    }

    Job &job_;
    zwz_ctx *ctx_;
    PinnedBuf in_, out_;
    std::string last_dir_;
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

// process.hpp:40. The reference runs one thread per archive (decompression.cpp:174); here the groups of all archives feed
// one worker pool.
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
    std::vector<Archive> archives;
    archives.reserve(names.size());
    for (const auto &n : names) {
        Archive a;
        a.filename = n;
        if (!map_archive(a)) {
            std::cerr << "Error opening file: " << n << std::endl;
            continue;
        }
        parse_archive(a.bytes, a.files);
        archives.push_back(std::move(a));
    }
    // groups of whole files, bounded by the raw bytes they may produce
    const uint64_t budget = std::max<uint64_t>(cfg.batch_bytes, 4 * CHUNK_SIZE);
    std::vector<Group> groups;
    for (size_t ai = 0; ai < archives.size(); ++ai) {
        const auto &files = archives[ai].files;
        size_t fi = 0;
        while (fi < files.size()) {
            if (files[fi].ordered.size() * CHUNK_SIZE > budget) {
                groups.push_back({ai, fi, 1, files[fi].ordered.size(), true});
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
            groups.push_back({ai, fi, fj - fi, nrec, false});
            fi = fj;
        }
    }
    stats().t_read += now_seconds() - t0;
    warm.join();
    if (warm_error) std::rethrow_exception(warm_error);

    Job job(archives, groups, output_dir, budget, cfg.device);
    const int workers = (int) std::min<size_t>((size_t) worker_count(), std::max<size_t>(1, groups.size()));
    if (!groups.empty())
        run_workers(workers, job.order, [&](int w) {
            Worker worker(job, w);
            worker.run();
        });
    for (auto &a : archives)
        if (a.map) munmap(a.map, a.bytes.n);
    print_timing("decompress");
}

} // namespace zwzhost
