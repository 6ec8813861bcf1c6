// decompress_pipeline.cpp — the decompress side (decompression.cpp:45-178 in the reference).
//
// The reference parses an archive record by record and inflates each one on the spot (one thread per archive, serial
// inside a file), holding early records in a per-path min-heap. Here an archive is parsed into a record index first
// (`.zwz` has no index of its own: records are self-delimiting, decompression.cpp:65-92), the reference's ordering rules
// are applied to that index, and whole files' worth of records go to the GPU in batches (zwz_decompress_records: one
// upload, batched inflate, device-side concatenation per file, MD5 of every output file from the same resident bytes,
// one download).
//
// Reader rules kept from the reference:
//   * a file receives the records with sequence ids 0, 1, 2, ... up to the first missing id (later ones would wait in
//     the heap forever, decompression.cpp:119-153, "Warning: pending chunks remaining"); of records sharing an id the
//     first one in archive order is used;
//   * a path gets its output file as soon as it is seen (decompression.cpp:95-110), even if nothing is ever written;
//   * the MD5 verdict line is printed only when the record carrying is_last_chunk arrived in order and nothing was
//     pending behind it (decompression.cpp:132) — ZWZ_VERIFY_ALL=1 prints a verdict for every complete file instead;
//   * errors inside a stream are not fatal: whatever inflate produced is written (decompression.cpp:31).
#include "zwz_host.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <stdexcept>

namespace zwzhost {

namespace fs = std::filesystem;
zwz_ctx *ctx_for(int device);
std::string md5_of_file_on(int device, const std::string &file_path);

namespace {

struct Rec {
    uint64_t off;   // payload offset inside the archive buffer
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

// grow-only page-locked host buffer (uploads/downloads straight from/to it run at full PCIe speed, and growing it does
// not zero-fill gigabytes the way std::vector::resize does)
struct PinnedBuf {
    zwz_ctx *ctx = nullptr;
    uint8_t *p = nullptr;
    size_t cap = 0, len = 0;
    explicit PinnedBuf(zwz_ctx *c) : ctx(c) {}
    ~PinnedBuf() {
        if (p) zwz_free_pinned(ctx, p);
    }
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) zwz_free_pinned(ctx, p);
        p = nullptr;
        cap = n + n / 8 + 4096;
        if (zwz_malloc_pinned(ctx, cap, (void **) &p) != ZWZ_OK) throw std::runtime_error("zwz: pinned allocation failed");
    }
    uint8_t *data() { return p; }
    const uint8_t *data() const { return p; }
    size_t size() const { return len; }
    uint8_t operator[](size_t i) const { return p[i]; }
};

bool read_whole(const std::string &path, PinnedBuf &buf) {
    std::FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.reserve(n > 0 ? (size_t) n : 1);
    buf.len = n > 0 ? std::fread(buf.data(), 1, (size_t) n, f) : 0;
    std::fclose(f);
    return true;
}

// decompression.cpp:65-92
void parse_archive(const PinnedBuf &a, std::vector<FileState> &files) {
    std::map<std::string, size_t> index;
    std::map<std::string, int> expected;                 // the reference's expected_sequence_id, simulated
    std::map<std::string, std::vector<int>> pending;     // seqs sitting in the heap
    size_t o = 0;
    const size_t n = a.size();
    while (o + 4 <= n) {
        int total_size, path_length;
        std::memcpy(&total_size, a.data() + o, 4);
        if (o + 8 > n) break;
        std::memcpy(&path_length, a.data() + o + 4, 4);
        if (path_length < 0 || o + 8 + (size_t) path_length + 5 > n) break;
        std::string relpath((const char *) a.data() + o + 8, (size_t) path_length);
        int sequence_id;
        std::memcpy(&sequence_id, a.data() + o + 8 + path_length, 4);
        bool is_last = a[o + 12 + path_length] != 0;
        long payload = (long) total_size - (4 + path_length + 4 + 1);
        size_t p0 = o + 13 + (size_t) path_length;
        if (payload < 0 || p0 + (size_t) payload > n) break;
        o = p0 + (size_t) payload;
        std::string md5;
        if (is_last) {
            if (o + MD5_DATA_SIZE > n) {
                md5.assign((const char *) a.data() + o, n - o);
                o = n;
            } else {
                md5.assign((const char *) a.data() + o, MD5_DATA_SIZE);
                o += MD5_DATA_SIZE;
            }
        }
        auto it = index.find(relpath);
        if (it == index.end()) {
            it = index.emplace(relpath, files.size()).first;
            files.emplace_back();
            files.back().relpath = relpath;
            expected[relpath] = 0;
        }
        FileState &fsx = files[it->second];
        fsx.recs.push_back({(uint64_t) p0, (uint32_t) payload, sequence_id, is_last});
        if (is_last) fsx.stored_md5 = md5;
        // simulate the heap to know whether the reference would print a verdict for this file
        int &exp = expected[relpath];
        auto &pend = pending[relpath];
        if (exp == sequence_id) {
            ++exp;
            for (;;) {
                auto m = std::min_element(pend.begin(), pend.end());
                if (m == pend.end() || *m != exp) break;
                pend.erase(m);
                ++exp;
            }
            if (is_last && exp == sequence_id + 1 && pend.empty()) fsx.verdict_in_order = true;
        } else {
            pend.push_back(sequence_id);
        }
    }
    for (auto &f : files) {
        int want = 0;
        for (;;) {
            const Rec *hit = nullptr;
            for (const auto &r : f.recs)
                if (r.seq == want) {
                    hit = &r;
                    break;
                }
            if (!hit) break;
            f.ordered.push_back(*hit);
            ++want;
        }
        f.complete = !f.ordered.empty() && f.ordered.back().last;
    }
}

void ensure_parent(const std::string &file_path) {
    fs::path dir = fs::path(file_path).parent_path();
    if (!dir.empty() && !fs::exists(dir)) {
        std::error_code ec;
        fs::create_directories(dir, ec);
    }
}

void print_verdict(const std::string &file_path, const std::string &stored, const std::string &calculated) {
    if (calculated != stored) { // decompression.cpp:140-146
        std::cerr << "MD5 mismatch for file: " << file_path << std::endl;
        std::cout << "Expected MD5: " << stored << std::endl;
        std::cout << "Calculated MD5: " << calculated << std::endl;
        stats().md5_mismatch++;
    } else {
        std::cout << "MD5 match for file: " << file_path << std::endl;
        stats().md5_match++;
    }
}

// A file whose records do not fit one batch: sub-batches of consecutive records are appended to the output file and the
// MD5 is taken the way the reference takes it — by reading the finished file back (decompression.cpp:136).
void big_file(zwz_ctx *ctx, const PinnedBuf &arch, FileState &fsx, const std::string &output_dir, uint64_t budget, PinnedBuf &out) {
    const RunConfig &cfg = config();
    std::string file_path = output_dir + "/" + fsx.relpath;
    ensure_parent(file_path);
    std::FILE *o = std::fopen(file_path.c_str(), "wb");
    if (!o) {
        std::cerr << "Error creating output file: " << file_path << std::endl;
        return;
    }
    const size_t per = std::max<size_t>(1, budget / CHUNK_SIZE);
    uint64_t written = 0;
    for (size_t r0 = 0; r0 < fsx.ordered.size(); r0 += per) {
        size_t r1 = std::min(fsx.ordered.size(), r0 + per), nrec = r1 - r0;
        std::vector<uint64_t> off(nrec), foff(2);
        std::vector<uint32_t> len(nrec), cap(nrec, (uint32_t) CHUNK_SIZE), rfile(nrec, 0u), raw_len(nrec), status(nrec);
        for (size_t i = 0; i < nrec; ++i) {
            off[i] = fsx.ordered[r0 + i].off;
            len[i] = fsx.ordered[r0 + i].len;
        }
        for (int attempt = 0;; ++attempt) {
            uint64_t need = 0;
            for (size_t i = 0; i < nrec; ++i) need += cap[i];
            out.reserve(need + 64);
            int rc = zwz_decompress_records(ctx, arch.data(), off.data(), len.data(), cap.data(), rfile.data(), (uint32_t) nrec, 1, out.data(),
                                            out.cap, foff.data(), raw_len.data(), status.data(), nullptr, 0);
            if (rc != ZWZ_OK) throw std::runtime_error(std::string("zwz_decompress_records: ") + zwz_last_error(ctx));
            bool again = false;
            for (size_t i = 0; i < nrec; ++i)
                if (status[i] == ZWZ_STREAM_OUTPUT_FULL) {
                    cap[i] = raw_len[i];
                    again = true;
                }
            if (!again || attempt >= 2) break;
        }
        if (foff[1]) std::fwrite(out.data(), 1, (size_t) foff[1], o);
        written += foff[1];
    }
    std::fclose(o);
    stats().files++;
    stats().records += fsx.ordered.size();
    stats().raw_bytes += written;
    if (fsx.recs.size() != fsx.ordered.size()) std::cerr << "Warning: pending chunks remaining for file: " << fsx.relpath << std::endl;
    if (fsx.complete && (fsx.verdict_in_order || cfg.verify_all)) print_verdict(file_path, fsx.stored_md5, md5_of_file_on(cfg.device, file_path));
}

void decompress_zwz(const std::string &filename, const std::string &output_dir) {
    const RunConfig &cfg = config();
    zwz_ctx *ctx = ctx_for(cfg.device);
    PinnedBuf arch(ctx);
    double t0 = now_seconds();
    if (!read_whole(filename, arch)) {
        std::cerr << "Error opening file: " << filename << std::endl;
        return;
    }
    std::vector<FileState> files;
    parse_archive(arch, files);
    stats().t_read += now_seconds() - t0;

    // groups of whole files, bounded by the raw bytes they may produce
    const uint64_t budget = std::max<uint64_t>(cfg.batch_bytes, 4 * CHUNK_SIZE);
    PinnedBuf out(ctx);
    size_t fi = 0;
    while (fi < files.size()) {
        if (files[fi].ordered.size() * CHUNK_SIZE > budget) { // one file larger than a batch: stream its records through
            big_file(ctx, arch, files[fi], output_dir, budget, out);
            ++fi;
            continue;
        }
        size_t fj = fi;
        uint64_t est = 0;
        size_t nrec = 0;
        while (fj < files.size() && (fj == fi || est + files[fj].ordered.size() * CHUNK_SIZE <= budget)) {
            est += files[fj].ordered.size() * CHUNK_SIZE;
            nrec += files[fj].ordered.size();
            ++fj;
        }
        const uint32_t nf = (uint32_t) (fj - fi);
        std::vector<uint64_t> off(nrec), foff(nf + 1);
        std::vector<uint32_t> len(nrec), cap(nrec, (uint32_t) CHUNK_SIZE), rfile(nrec), raw_len(nrec), status(nrec);
        size_t k = 0;
        for (size_t f = fi; f < fj; ++f)
            for (const auto &r : files[f].ordered) {
                off[k] = r.off;
                len[k] = r.len;
                rfile[k] = (uint32_t) (f - fi);
                ++k;
            }
        std::vector<uint8_t> digest((size_t) nf * 16);
        t0 = now_seconds();
        for (int attempt = 0;; ++attempt) {
            uint64_t need = 0;
            for (size_t i = 0; i < nrec; ++i) need += cap[i];
            out.reserve(need + 64);
            int rc = zwz_decompress_records(ctx, arch.data(), off.data(), len.data(), cap.data(), rfile.data(), (uint32_t) nrec, nf, out.data(),
                                            out.cap, foff.data(), raw_len.data(), status.data(), digest.data(), 0);
            if (rc != ZWZ_OK) throw std::runtime_error(std::string("zwz_decompress_records: ") + zwz_last_error(ctx));
            bool again = false;
            for (size_t i = 0; i < nrec; ++i)
                if (status[i] == ZWZ_STREAM_OUTPUT_FULL) { // a foreign record larger than 65 535 bytes: the reference's loop handles any size
                    cap[i] = raw_len[i];
                    again = true;
                }
            if (!again || attempt >= 2) break;
        }
        stats().t_gpu += now_seconds() - t0;
        t0 = now_seconds();
        for (size_t f = fi; f < fj; ++f) {
            FileState &fsx = files[f];
            std::string file_path = output_dir + "/" + fsx.relpath;
            ensure_parent(file_path);
            std::FILE *o = std::fopen(file_path.c_str(), "wb");
            if (!o) {
                std::cerr << "Error creating output file: " << file_path << std::endl;
                continue;
            }
            uint64_t a0 = foff[f - fi], a1 = foff[f - fi + 1];
            if (a1 > a0) std::fwrite(out.data() + a0, 1, (size_t) (a1 - a0), o);
            std::fclose(o);
            stats().files++;
            stats().records += fsx.ordered.size();
            stats().raw_bytes += a1 - a0;
            if (fsx.recs.size() != fsx.ordered.size()) std::cerr << "Warning: pending chunks remaining for file: " << fsx.relpath << std::endl;
            if (cfg.verbose && !fsx.stored_md5.empty()) std::cout << "Read MD5: " << fsx.stored_md5 << std::endl; // decompression.cpp:90
            if (fsx.complete && (fsx.verdict_in_order || cfg.verify_all)) {
                char hex[32];
                zwz_md5_hex(&digest[(f - fi) * 16], hex);
                print_verdict(file_path, fsx.stored_md5, std::string(hex, 32));
            }
        }
        stats().t_write += now_seconds() - t0;
        fi = fj;
    }
}

} // namespace

// process.hpp:40. Archives are independent; they are taken one after the other here because every one of them already
// fills the GPU (the reference's parallelism across archives, decompression.cpp:174, was its only parallelism).
void do_decompression(const std::string &input_dir, const std::string &output_dir) {
    std::vector<std::string> archives;
    for (const auto &entry : fs::directory_iterator(input_dir)) {
        if (entry.path().extension() == ".zwz") archives.push_back(entry.path().string()); // decompression.cpp:168-172
    }
    std::sort(archives.begin(), archives.end());
    for (const auto &a : archives) decompress_zwz(a, output_dir);
    print_timing("decompress");
}

} // namespace zwzhost
