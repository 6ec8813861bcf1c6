// compress_pipeline.cpp — the compress side of one rank (compression.cpp:24-194 in the reference).
//
// The reference runs a producer thread that cuts files into 65 535-byte chunks, a mutex queue, ONE consumer thread that
// calls zlib per chunk, and a writer that re-reads each source file for its MD5. Here the rank's files are planned into
// batches up front — whole files per batch, and files larger than a batch cut into chunk-aligned SEGMENTS that are batches
// of their own — and W workers (pipeline.hpp) each take a batch: read it into their page-locked staging buffer (several
// reader threads per batch: 370 000 small files are open/read/close bound), make ONE trip through the GPU, serialise the
// records in the reference's byte layout (compression.cpp:73-104) and take the batch's place in the archive. Offsets are
// handed out in batch order, so the archive is the same bytes for any W: record order is file order then sequence order,
// which is exactly what the reference produces with NUM_CONSUMERS = 1.
//
// MD5 is off the batch critical path. A file's digest is one serial chain (~0.13 GB/s per file), so:
//   * batches of whole files go through zwz_compress_files_async: the call returns when the packed streams are back, the
//     digests follow on the context's second stream; the serialised records wait in memory (two batches per worker) until
//     zwz_wait delivers the digests, then go to the archive with ONE pwrite at the offset reserved earlier;
//   * a file cut into segments is hashed by a separate hasher thread that streams it through zwz_md5_update_device on its
//     own context while all workers deflate its segments; only the record that carries the digest waits for it.
//
// One file over several GPUs (BASELINE config 3: a single 16 GB file). The reference's deal is file-granular
// (compression.cpp:31-41), so such a file lands on one rank and one deflate thread, and its reader wants all records of a
// path in ONE archive (decompression.cpp:52-55). Here the segments of a file that is cut are dealt over the ranks of the run:
// the file's owner (the rank the deal gives it to) keeps the archive order and hands every segment its place — archive
// offset and first sequence id — as soon as the rank that deflates it has published how many bytes and records it made
// (the segment ledger, pipeline.hpp); that rank then writes its records straight into the owner's archive with pwrite. The
// exchange is two 16-byte entries per segment; the payload never crosses ranks.
#include "pipeline.hpp"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <future>
#include <iostream>
#include <sstream>
#include <unistd.h>

namespace zwzhost {

namespace fs = std::filesystem;

namespace {

struct PlannedFile {
    std::string relpath;
    uint64_t size = 0; // from stat at planning time; the bytes actually read decide the records
    int big = -1;      // index into Job::bigs when the file is cut into segments
};
struct Batch {
    size_t first = 0, count = 0; // files [first, first + count); a segment: count == 1
    uint64_t bytes = 0;
    bool segment = false;
    uint64_t seg_off = 0, seg_len = 0; // byte range of the file; a whole number of chunks unless it is the last segment
    bool seg_last = false;
    int seg = 0;               // segment number inside the file
    // who does what with a segment when the file is cut over several ranks: LOCAL = this rank deflates it and owns the archive;
    // PLACE = this rank owns the archive, another rank deflates (the owner only hands out the place); REMOTE = this rank
    // deflates for another rank's archive
    enum Role { LOCAL, PLACE, REMOTE } role = LOCAL;
};
struct LoadedFile {
    const std::string *relpath;
    uint64_t off = 0, size = 0; // inside the staging buffer
};
// a file that is cut into segments: its sequence ids continue across batches (the split rule may add records), and its
// digest comes from the hasher thread
struct BigFile {
    size_t file = 0;                     // index into the rank's file list
    int owner = 0;                       // rank whose archive holds the file
    std::string key;                     // ledger key (same on every rank)
    int next_seq = 0;                    // guarded by the commit order (owner only)
    std::shared_future<std::string> md5; // 32 hex characters ("" when the file could not be read); owner only
};

// compression.cpp:73-104: i32 total_size, i32 path_len, path, i32 sequence_id, u8 is_last_chunk, payload[, 32 hex chars]
void append_record(std::vector<char> &out, const std::string &relpath, int sequence_id, bool is_last_chunk, const uint8_t *payload,
                   uint32_t payload_len, RunStats &st) {
    int path_length = static_cast<int>(relpath.size());
    int total_size = (int) sizeof(path_length) + path_length + (int) sizeof(sequence_id) + (int) sizeof(bool) + (int) payload_len;
    size_t o = out.size();
    out.resize(o + 4 + 4 + relpath.size() + 4 + 1 + payload_len + (is_last_chunk ? MD5_DATA_SIZE : 0));
    char *p = out.data() + o;
    std::memcpy(p, &total_size, 4);
    std::memcpy(p + 4, &path_length, 4);
    std::memcpy(p + 8, relpath.data(), relpath.size());
    p += 8 + relpath.size();
    std::memcpy(p, &sequence_id, 4);
    p[4] = is_last_chunk ? 1 : 0;
    std::memcpy(p + 5, payload, payload_len);
    if (is_last_chunk) std::memset(p + 5 + payload_len, '0', MD5_DATA_SIZE); // filled in when the digest arrives
    st.records++;
    st.payload_bytes += payload_len;
}

// one chunk -> one record, or two when the split rule fired (the reference would have silently truncated this chunk)
void emit_chunk(std::vector<char> &out, const std::string &relpath, int &seq, bool last_chunk_of_file, const uint8_t *payload,
                const zwz_deflate_result &r, RunStats &st) {
    if (r.len1 == 0) {
        append_record(out, relpath, seq++, last_chunk_of_file, payload, r.len0, st);
    } else {
        append_record(out, relpath, seq++, false, payload, r.len0, st);
        append_record(out, relpath, seq++, last_chunk_of_file, payload + r.len0, r.len1, st);
    }
}

void write_at(int fd, const char *p, size_t n, uint64_t off) {
    while (n) {
        ssize_t w = ::pwrite(fd, p, n, (off_t) off);
        if (w < 0) {
            if (errno == EINTR) continue;
            throw std::runtime_error("zwz: short write to archive");
        }
        p += w;
        n -= (size_t) w;
        off += (uint64_t) w;
    }
}

std::string generate_output_filename(const std::string &output_dir, int world_rank) { // compression.cpp:151-159
    fs::path dir(output_dir);
    dir = dir / "";
    return (dir / ("compressed_" + std::to_string(world_rank) + ".zwz")).string();
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
    g.payload_bytes += s.payload_bytes;
    g.t_read += s.t_read;
    g.t_gpu += s.t_gpu;
    g.t_write += s.t_write;
}

// shared by the workers of one do_compression call
struct Job {
    const std::string &input_dir;
    const std::string &output_dir;
    const std::vector<PlannedFile> &files;
    const std::vector<Batch> &batches;
    std::vector<BigFile> &bigs;
    int fd;
    size_t cap;     // staging capacity of a worker
    size_t out_cap; // enough for the packed streams of any planned batch: page-locked buffers are sized once per worker
    int level;
    int device;
    std::atomic<size_t> next_batch{0};
    OrderedCommit order;
    uint64_t archive_off = 0; // guarded by the commit order
    Job(const std::string &in, const std::string &out, const std::vector<PlannedFile> &f, const std::vector<Batch> &b, std::vector<BigFile> &bg, int fd_,
        size_t cap_, int level_, int device_)
        : input_dir(in), output_dir(out), files(f), batches(b), bigs(bg), fd(fd_), cap(cap_), out_cap(0), level(level_), device(device_) {
        size_t max_files = 1;
        for (const auto &x : b) {
            max_files = std::max(max_files, x.count);
            cap = std::max<size_t>(cap, (size_t) x.bytes);
        }
        out_cap = cap + (cap / CHUNK_SIZE + max_files + 16) * 64 + 4096;
    }
};

// the digest of a file that is cut into segments: the whole file streamed through the GPU's MD5 update/final pair on the
// hasher's own context (verification.cpp:13-22 is the same loop with 1 024-byte pieces on the CPU)
std::string hash_big_file(int device, int hasher, const std::string &full) {
    return md5_of_file_ctx(worker_ctx(device, 64 + hasher), full);
}

class Worker {
  public:
    Worker(Job &job, int id) : job_(job), ctx_(worker_ctx(job.device, id)), stage_(ctx_), out_(ctx_) {}

    void run() {
        for (;;) {
            Pending &p = pend_[turn_++ & 1];
            finalize(p); // its buffers are about to be reused. BEFORE a batch index is taken: waiting for digests while holding
                         // an index would keep every later batch from committing
            size_t b = job_.next_batch.fetch_add(1);
            if (b >= job_.batches.size()) break;
            const Batch &batch = job_.batches[b];
            if (batch.segment && batch.role == Batch::PLACE)
                place(b, batch, p);
            else if (batch.segment)
                segment(b, batch, p);
            else
                small_files(b, batch, p);
        }
        finalize(pend_[turn_ & 1]);
        finalize(pend_[(turn_ + 1) & 1]);
    }

  private:
    // a batch whose records are serialised and whose place in the archive is reserved, waiting for its digests
    struct Pending {
        bool active = false;
        uint64_t ticket = 0;
        std::vector<char> records;
        uint64_t archive_off = 0;
        std::vector<size_t> md5_pos;   // per file of the batch: where its 32 hex characters go inside `records`
        std::vector<uint8_t> digest;   // 16 bytes per file, filled by zwz_wait
        std::shared_future<std::string> big_md5; // a segment that carries its file's digest
        bool has_big_md5 = false;
        RunStats st;
    };

    void finalize(Pending &p) {
        if (!p.active) return;
        p.active = false;
        double t0 = now_seconds();
        if (p.ticket && zwz_wait(ctx_, p.ticket) != ZWZ_OK) throw std::runtime_error(std::string("zwz_wait: ") + zwz_last_error(ctx_));
        for (size_t i = 0; i < p.md5_pos.size(); ++i) {
            if (p.has_big_md5) {
                const std::string hex = p.big_md5.get();
                // verification.cpp:8-11: an unreadable file gives "" and NOTHING is written — here the field stays zeros,
                // which keeps the archive parseable
                if (hex.size() == MD5_DATA_SIZE) std::memcpy(p.records.data() + p.md5_pos[i], hex.data(), MD5_DATA_SIZE);
            } else {
                zwz_md5_hex(&p.digest[i * 16], p.records.data() + p.md5_pos[i]);
            }
        }
        p.st.t_gpu += now_seconds() - t0;
        t0 = now_seconds();
        write_at(job_.fd, p.records.data(), p.records.size(), p.archive_off);
        p.st.t_write += now_seconds() - t0;
        merge_stats(p.st);
        p.st = RunStats();
    }

    // takes this batch's place in the archive, in batch order; the write itself runs unordered, later
    template <class F> void commit(size_t b, Pending &p, const std::string &log_text, F &&ordered_fixup) {
        job_.order.wait_turn(b);
        ordered_fixup();
        p.archive_off = job_.archive_off;
        job_.archive_off += p.records.size();
        if (!log_text.empty()) std::cerr << log_text << std::flush;
        job_.order.done(b);
        p.active = true;
    }

    void small_files(size_t b, const Batch &batch, Pending &p) {
        RunStats &st = p.st;
        std::ostringstream log;
        stage_.reserve(job_.cap + 64);
        // ---- read: the files' places in the staging buffer are known from the plan, so readers work independently
        double t0 = now_seconds();
        const size_t nfiles = batch.count;
        std::vector<uint64_t> at(nfiles + 1, 0), got(nfiles, 0);
        std::vector<char> opened(nfiles, 0);
        for (size_t i = 0; i < nfiles; ++i) at[i + 1] = at[i] + job_.files[batch.first + i].size;
        parallel_for(nfiles, io_threads(), [&](size_t i) {
            const PlannedFile &pf = job_.files[batch.first + i];
            std::string full = job_.input_dir + "/" + pf.relpath;
            int fd = ::open(full.c_str(), O_RDONLY);
            if (fd < 0) return;
            opened[i] = 1;
            uint64_t done = 0;
            while (done < pf.size) {
                ssize_t r = ::read(fd, stage_.data() + at[i] + done, pf.size - done);
                if (r < 0 && errno == EINTR) continue;
                if (r <= 0) break;
                done += (uint64_t) r;
            }
            ::close(fd);
            got[i] = done;
        });
        // compact (a file that shrank or vanished since the plan leaves a hole)
        std::vector<LoadedFile> loaded;
        loaded.reserve(nfiles);
        uint64_t used = 0;
        for (size_t i = 0; i < nfiles; ++i) {
            const PlannedFile &pf = job_.files[batch.first + i];
            if (!opened[i]) { // compression.cpp:45-48: report and skip
                log << "Error opening source file: " << job_.input_dir << "/" << pf.relpath << "\n";
                continue;
            }
            if (used != at[i] && got[i]) std::memmove(stage_.data() + used, stage_.data() + at[i], got[i]);
            loaded.push_back({&pf.relpath, used, got[i]});
            used += got[i];
        }
        st.t_read += now_seconds() - t0;
        p.records.clear();
        p.md5_pos.clear();
        p.ticket = 0;
        p.has_big_md5 = false;
        if (!loaded.empty()) {
            // ---- GPU
            const uint32_t nf = (uint32_t) loaded.size();
            std::vector<uint64_t> foff(nf + 1);
            for (uint32_t i = 0; i < nf; ++i) foff[i] = loaded[i].off;
            foff[nf] = used;
            const uint64_t nc = zwz_count_chunks(foff.data(), nf);
            std::vector<uint64_t> poff(nc + 1);
            std::vector<zwz_deflate_result> res(nc);
            p.digest.assign((size_t) nf * 16, 0);
            out_.reserve(std::max<size_t>(job_.out_cap, used + nc * 64 + 4096));
            t0 = now_seconds();
            int rc = zwz_compress_files_async(ctx_, stage_.data(), foff.data(), nf, job_.level, out_.data(), out_.cap, poff.data(), res.data(),
                                              p.digest.data(), &p.ticket);
            if (rc != ZWZ_OK) throw std::runtime_error(std::string("zwz_compress_files: ") + zwz_last_error(ctx_));
            st.t_gpu += now_seconds() - t0;
            // ---- serialise
            t0 = now_seconds();
            p.records.reserve((size_t) poff[nc] + nc * 24 + (size_t) nf * 96);
            size_t c = 0;
            for (uint32_t i = 0; i < nf; ++i) {
                uint64_t nch = loaded[i].size / CHUNK_SIZE + 1;
                int seq = 0;
                for (uint64_t k = 0; k < nch; ++k, ++c) emit_chunk(p.records, *loaded[i].relpath, seq, k + 1 == nch, out_.data() + poff[c], res[c], st);
                p.md5_pos.push_back(p.records.size() - MD5_DATA_SIZE);
                st.files++;
                st.raw_bytes += loaded[i].size;
                if (config().verbose) log << "md5 value size: " << MD5_DATA_SIZE << "\n"; // compression.cpp:100
            }
            st.t_write += now_seconds() - t0;
        }
        commit(b, p, log.str(), [] {});
    }

    // One segment of a file that does not fit a batch: k x 65 535 bytes (the last one: the rest, with the possibly empty tail
    // chunk). Deflate only — the file's digest comes from the hasher thread.
    void segment(size_t b, const Batch &batch, Pending &p) {
        RunStats &st = p.st;
        const PlannedFile &pf = job_.files[batch.first];
        BigFile &big = job_.bigs[(size_t) pf.big];
        const std::string full = job_.input_dir + "/" + pf.relpath;
        stage_.reserve(job_.cap + 64);
        p.records.clear();
        p.md5_pos.clear();
        p.ticket = 0;
        p.has_big_md5 = false;
        std::string log;
        double t0 = now_seconds();
        int fd = ::open(full.c_str(), O_RDONLY);
        std::atomic<uint64_t> got{0};
        if (fd >= 0) {
            const uint64_t piece = (uint64_t) 8 << 20;
            const size_t npieces = (size_t) ((batch.seg_len + piece - 1) / piece);
            parallel_for(npieces, io_threads(), [&](size_t i) {
                uint64_t o = i * piece, n = std::min<uint64_t>(piece, batch.seg_len - o), done = 0;
                while (done < n) {
                    ssize_t r = ::pread(fd, stage_.data() + o + done, n - done, (off_t) (batch.seg_off + o + done));
                    if (r < 0 && errno == EINTR) continue;
                    if (r <= 0) break;
                    done += (uint64_t) r;
                }
                got += done;
            });
            ::close(fd);
        } else if (batch.seg_off == 0) {
            log = "Error opening source file: " + full + "\n";
        }
        st.t_read += now_seconds() - t0;
        const bool usable = fd >= 0 && got.load() == batch.seg_len; // a file that changed under us: its remaining records are dropped
        int provisional = (int) (batch.seg_off / CHUNK_SIZE);
        if (usable) {
            const uint64_t nfull = batch.seg_len / CHUNK_SIZE;
            const uint64_t nch = batch.seg_last ? nfull + 1 : nfull;
            std::vector<uint64_t> off(nch), poff(nch + 1);
            std::vector<uint32_t> len(nch);
            std::vector<zwz_deflate_result> res(nch);
            for (uint64_t k = 0; k < nch; ++k) {
                off[k] = k * CHUNK_SIZE;
                len[k] = (uint32_t) std::min<uint64_t>(CHUNK_SIZE, batch.seg_len - k * CHUNK_SIZE);
            }
            out_.reserve(std::max<size_t>(job_.out_cap, batch.seg_len + nch * 64 + 4096));
            t0 = now_seconds();
            if (nch && zwz_deflate_batch(ctx_, stage_.data(), off.data(), len.data(), (uint32_t) nch, out_.data(), out_.cap, poff.data(), res.data(),
                                         job_.level) != ZWZ_OK)
                throw std::runtime_error(std::string("zwz_deflate_batch: ") + zwz_last_error(ctx_));
            st.t_gpu += now_seconds() - t0;
            t0 = now_seconds();
            p.records.reserve((size_t) poff[nch] + nch * (24 + pf.relpath.size()) + 64);
            int seq = provisional;
            for (uint64_t k = 0; k < nch; ++k) emit_chunk(p.records, pf.relpath, seq, batch.seg_last && k + 1 == nch, out_.data() + poff[k], res[k], st);
            if (batch.seg_last) {
                p.md5_pos.push_back(p.records.size() - MD5_DATA_SIZE);
                p.big_md5 = big.md5;
                p.has_big_md5 = true;
                st.files++;
            }
            st.raw_bytes += batch.seg_len;
            st.t_write += now_seconds() - t0;
        }
        if (batch.role == Batch::REMOTE) {
            // another rank's archive: say what this segment came to, learn its place, write it there. Nothing of this enters
            // this rank's own archive, so its turn in the commit order is passed on at once.
            job_.order.wait_turn(b);
            if (!log.empty()) std::cerr << log << std::flush;
            job_.order.done(b);
            uint32_t nrec = 0;
            for (size_t o = 0; o + 8 <= p.records.size();) {
                int total_size, path_length;
                std::memcpy(&total_size, p.records.data() + o, 4);
                std::memcpy(&path_length, p.records.data() + o + 4, 4);
                const bool last = p.records[o + 12 + (size_t) path_length] != 0;
                o += 4 + (size_t) total_size + (last ? MD5_DATA_SIZE : 0);
                ++nrec;
            }
            ledger_publish(job_.output_dir, big.key, batch.seg, ".size", p.records.size(), nrec);
            uint64_t seq_base = 0;
            const uint64_t at = ledger_wait(job_.output_dir, big.key, batch.seg, ".place", &seq_base);
            renumber(p.records, (int) seq_base);
            const std::string owner_archive = generate_output_filename(job_.output_dir, big.owner);
            int ofd = ::open(owner_archive.c_str(), O_WRONLY);
            if (ofd < 0) throw std::runtime_error("zwz: cannot open " + owner_archive);
            double t1 = now_seconds();
            write_at(ofd, p.records.data(), p.records.size(), at);
            ::close(ofd);
            st.t_write += now_seconds() - t1;
            ledger_publish(job_.output_dir, big.key, batch.seg, ".done", p.records.size(), nrec);
            merge_stats(st);
            p.st = RunStats();
            p.records.clear();
            return; // nothing pending
        }
        commit(b, p, log, [&] {
            // sequence ids continue where the previous segment stopped (they run ahead of the chunk index when the split rule
            // fired earlier in the file): renumber in place, the id sits at a fixed place of every record header
            big.next_seq += renumber(p.records, big.next_seq);
        });
    }

    // The owner's side of a segment another rank deflates: wait for its size, hand out its place, move on.
    void place(size_t b, const Batch &batch, Pending &p) {
        const PlannedFile &pf = job_.files[batch.first];
        BigFile &big = job_.bigs[(size_t) pf.big];
        p.records.clear();
        p.md5_pos.clear();
        p.ticket = 0;
        p.has_big_md5 = false;
        job_.order.wait_turn(b);
        uint64_t nrec = 0;
        const uint64_t bytes = ledger_wait(job_.output_dir, big.key, batch.seg, ".size", &nrec);
        ledger_publish(job_.output_dir, big.key, batch.seg, ".place", job_.archive_off, (uint64_t) big.next_seq);
        job_.archive_off += bytes;
        big.next_seq += (int) nrec;
        job_.order.done(b);
        ledger_wait(job_.output_dir, big.key, batch.seg, ".done"); // the archive is complete only when the other rank has written
        p.st.records += nrec;
        p.st.raw_bytes += batch.seg_len;
        merge_stats(p.st);
        p.st = RunStats();
    }

    // sets the sequence ids of the serialised records to base, base + 1, ...; returns how many there are
    static int renumber(std::vector<char> &records, int base) {
        int count = 0;
        size_t o = 0;
        while (o + 8 <= records.size()) {
            int total_size, path_length;
            std::memcpy(&total_size, records.data() + o, 4);
            std::memcpy(&path_length, records.data() + o + 4, 4);
            int seq = base + count++;
            std::memcpy(records.data() + o + 8 + path_length, &seq, 4);
            const bool last = records[o + 12 + (size_t) path_length] != 0;
            o += 4 + (size_t) total_size + (last ? MD5_DATA_SIZE : 0);
        }
        return count;
    }

    Job &job_;
    zwz_ctx *ctx_;
    PinnedBuf stage_, out_;
    Pending pend_[2];
    unsigned turn_ = 0;
};

} // namespace

// Same signature and meaning as process.hpp:39. The rank takes lines world_rank, world_rank + P, ... of the record file
// (compression.cpp:31-41; every line counts, blank or not, exactly like the getline loop there).
void do_compression(const std::string &input_dir, const std::string &output_dir, const std::string &file_record, int world_rank) {
    const RunConfig &cfg = config();
    std::ifstream record_file(file_record);
    if (!record_file.is_open()) {
        std::cerr << "Rank: " << world_rank << " - Error opening file record: " << file_record << std::endl;
        return;
    }
    std::string output_filename = generate_output_filename(output_dir, world_rank);
    const int line_count = count_non_empty_lines(file_record);
    // like the reference (main.cpp:47-52), a rank beyond the number of files writes no archive — but it may still deflate
    // segments of a big file for another rank's archive
    const bool has_archive = world_rank < line_count;
    int fd = -1;
    if (has_archive) {
        fd = ::open(output_filename.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) {
            std::cerr << "Rank: " << world_rank << " - Error opening archive: " << output_filename << std::endl;
            return;
        }
        std::cout << "Max record line num: " << line_count << std::endl;
    }

    // the device comes up while the files are dealt and sized
    std::exception_ptr warm_error;
    std::thread warm([&] {
        try {
            ctx_for(cfg.device);
        } catch (...) {
            warm_error = std::current_exception();
        }
    });
    // the deal (compression.cpp:31-41), then the plan, in deal order
    std::vector<std::string> lines;
    {
        std::string file_path;
        while (std::getline(record_file, file_path)) lines.push_back(file_path);
    }
    const int file_number = (int) lines.size();
    const int world = std::max(1, cfg.world_size);
    std::vector<PlannedFile> files;
    for (size_t i = (size_t) world_rank; i < lines.size(); i += (size_t) world) files.push_back({lines[i], 0, -1});
    const size_t n_own = files.size();
    auto size_of = [&](const std::string &rel) {
        std::error_code ec;
        uint64_t size = fs::file_size(fs::path(input_dir) / rel, ec);
        return ec ? (uint64_t) 0 : size;
    };
    parallel_for(n_own, io_threads() * 2, [&](size_t i) { files[i].size = size_of(files[i].relpath); });
    uint64_t total_bytes = 0;
    for (const auto &f : files) total_bytes += f.size;
    // batches small enough that every worker gets several (a 200 MB job in one 128 MB batch would leave the pool idle)
    const int pool = worker_count();
    const size_t cap = (size_t) std::min<uint64_t>(cfg.batch_bytes, std::max<uint64_t>((uint64_t) 8 << 20, total_bytes / ((uint64_t) pool * 3) + 1));
    // Files larger than a batch (cfg.batch_bytes: the same on every rank) are cut into segments of whole chunks. The record
    // is size-descending, so the run's big files are its first lines: every rank finds the same list.
    const uint64_t seg_bytes = std::max<uint64_t>(1, cfg.batch_bytes / CHUNK_SIZE) * CHUNK_SIZE;
    struct RunBig {
        size_t line;
        uint64_t size;
    };
    std::vector<RunBig> run_bigs;
    for (size_t gi = 0; gi < lines.size(); ++gi) {
        const uint64_t size = (int) (gi % (size_t) world) == world_rank ? files[gi / (size_t) world].size : (world > 1 ? size_of(lines[gi]) : 0);
        if (size + 64 <= cfg.batch_bytes) break;
        run_bigs.push_back({gi, size});
    }
    std::vector<Batch> batches;
    std::vector<BigFile> bigs;
    auto segments_of = [&](size_t file_index, int big_index, uint64_t size, int owner, bool mine_is_owner) {
        const int nseg = (int) (size / seg_bytes) + 1; // the last segment carries the tail chunk (possibly empty)
        for (int sgi = 0; sgi < nseg; ++sgi) {
            const bool last = sgi == nseg - 1;
            const int worker_rank = (last || world == 1) ? owner : (owner + sgi) % world; // the digest travels with the last one
            Batch b;
            b.first = file_index;
            b.count = 1;
            b.segment = true;
            b.seg = sgi;
            b.seg_off = (uint64_t) sgi * seg_bytes;
            b.seg_len = last ? size - b.seg_off : seg_bytes;
            b.bytes = b.seg_len;
            b.seg_last = last;
            if (mine_is_owner)
                b.role = worker_rank == world_rank ? Batch::LOCAL : Batch::PLACE;
            else if (worker_rank == world_rank)
                b.role = Batch::REMOTE;
            else
                continue;
            if (b.role == Batch::PLACE) b.bytes = 0;
            batches.push_back(b);
        }
        (void) big_index;
    };
    // segments this rank deflates for other ranks' archives come first: their owners are waiting for the sizes
    for (const auto &rb : run_bigs) {
        const int owner = (int) (rb.line % (size_t) world);
        if (owner == world_rank) continue;
        files.push_back({lines[rb.line], rb.size, (int) bigs.size()});
        BigFile bf;
        bf.file = files.size() - 1;
        bf.owner = owner;
        bf.key = "c" + std::to_string(rb.line) + "_" + cfg.run_id;
        bigs.push_back(bf);
        segments_of(files.size() - 1, (int) bigs.size() - 1, rb.size, owner, false);
    }
    for (size_t i = 0; i < n_own; ++i) {
        if (files[i].size + 64 > cfg.batch_bytes) {
            files[i].big = (int) bigs.size();
            BigFile bf;
            bf.file = i;
            bf.owner = world_rank;
            bf.key = "c" + std::to_string((size_t) world_rank + i * (size_t) world) + "_" + cfg.run_id;
            bigs.push_back(bf);
            segments_of(i, files[i].big, files[i].size, world_rank, true);
            continue;
        }
        if (batches.empty() || batches.back().segment || batches.back().bytes + files[i].size > cap || batches.back().count >= 262144) {
            Batch nb;
            nb.first = i;
            batches.push_back(nb);
        }
        batches.back().count++;
        batches.back().bytes += files[i].size;
    }
    if (world > 1 && !run_bigs.empty()) {
        std::error_code ec;
        fs::create_directories(ledger_dir(output_dir), ec);
    }

    if (batches.empty() && !has_archive) { // nothing of its own and nothing to do for the others
        warm.join();
        std::cout << "Rank: " << world_rank << " - No file to compress" << std::endl;
        return;
    }
    warm.join();
    if (warm_error) {
        if (fd >= 0) ::close(fd);
        std::rethrow_exception(warm_error);
    }
    // hashers for the files that are cut into segments (at most 4 at a time; each streams its file once more through the GPU)
    std::vector<std::thread> hashers;
    std::vector<std::promise<std::string>> promises(bigs.size());
    for (size_t k = 0; k < bigs.size(); ++k) bigs[k].md5 = promises[k].get_future().share();
    std::atomic<size_t> next_big{0};
    for (size_t h = 0; h < std::min<size_t>(bigs.size(), 4); ++h)
        hashers.emplace_back([&, h] {
            for (;;) {
                size_t k = next_big.fetch_add(1);
                if (k >= bigs.size()) return;
                if (bigs[k].owner != world_rank) { // another rank's file: its owner hashes it
                    promises[k].set_value("");
                    continue;
                }
                try {
                    promises[k].set_value(hash_big_file(cfg.device, (int) h, input_dir + "/" + files[bigs[k].file].relpath));
                } catch (...) {
                    promises[k].set_exception(std::current_exception());
                }
            }
        });
    Job job(input_dir, output_dir, files, batches, bigs, fd, cap, cfg.level, cfg.device);
    const int workers = (int) std::min<size_t>((size_t) pool, std::max<size_t>(1, batches.size()));
    std::exception_ptr failure;
    try {
        run_workers(workers, job.order, [&](int w) {
            Worker worker(job, w);
            worker.run();
        });
    } catch (...) {
        failure = std::current_exception();
    }
    for (auto &t : hashers) t.join();
    if (fd >= 0) ::close(fd);
    if (failure) std::rethrow_exception(failure);
    if (world > 1) { // this rank's ledger entries have served their purpose (every PLACE waited for its segment's .done)
        std::error_code ec;
        for (const auto &bf : bigs) {
            if (bf.owner != world_rank) continue;
            const int nseg = (int) (files[bf.file].size / seg_bytes) + 1;
            for (int sg = 0; sg < nseg; ++sg)
                for (const char *what : {".size", ".place", ".done"}) fs::remove(ledger_dir(output_dir) + "/" + bf.key + "." + std::to_string(sg) + what, ec);
        }
        fs::remove(ledger_dir(output_dir), ec); // succeeds once the last owner has cleaned up
    }
    if (!has_archive) {
        std::cout << "Rank: " << world_rank << " - No file to compress" << std::endl;
        return;
    }
    std::cout << "Rank: " << world_rank << " - Total processed file: " << file_number << std::endl;
    print_timing("compress");
}

} // namespace zwzhost
