// compress_pipeline.cpp — the compress side of one rank (compression.cpp:24-194 in the reference).
//
// The reference runs a producer thread that cuts files into 65 535-byte chunks, a mutex queue, ONE consumer thread that
// calls zlib per chunk, and a writer that re-reads each source file for its MD5. Here the rank's files are planned into
// batches of whole files up front; W workers (pipeline.hpp) each take a batch, read its files into their page-locked staging
// buffer, make ONE trip through the GPU (zwz_compress_files: upload once, deflate every chunk, MD5 every file from the same
// resident bytes, pack, download), serialise the records in the reference's byte layout (compression.cpp:73-104) and write
// them at the batch's offset of the archive. Offsets are handed out in batch order, so the archive is the same bytes for
// any W: record order is file order then sequence order, which is exactly what the reference produces with
// NUM_CONSUMERS = 1.
#include "pipeline.hpp"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <iostream>
#include <sstream>
#include <unistd.h>

namespace zwzhost {

namespace fs = std::filesystem;

namespace {

struct PlannedFile {
    std::string relpath;
    uint64_t size = 0; // from stat at planning time; the bytes actually read decide the records
};
struct Batch {
    size_t first = 0, count = 0;
    uint64_t bytes = 0;
    bool big = false; // one file larger than the staging buffer: streamed through in segments
};
struct LoadedFile {
    const std::string *relpath;
    uint64_t off = 0, size = 0; // inside the staging buffer
};

// compression.cpp:73-104: i32 total_size, i32 path_len, path, i32 sequence_id, u8 is_last_chunk, payload[, 32 hex chars]
void append_record(std::vector<char> &out, const std::string &relpath, int sequence_id, bool is_last_chunk, const uint8_t *payload,
                   uint32_t payload_len, const char *md5_hex, RunStats &st) {
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
    if (is_last_chunk) std::memcpy(p + 5 + payload_len, md5_hex, MD5_DATA_SIZE);
    st.records++;
    st.payload_bytes += payload_len;
}

// one chunk -> one record, or two when the split rule fired (the reference would have silently truncated this chunk)
void emit_chunk(std::vector<char> &out, const std::string &relpath, int &seq, bool last_chunk_of_file, const uint8_t *payload,
                const zwz_deflate_result &r, const char *md5_hex, RunStats &st) {
    if (r.len1 == 0) {
        append_record(out, relpath, seq++, last_chunk_of_file, payload, r.len0, md5_hex, st);
    } else {
        append_record(out, relpath, seq++, false, payload, r.len0, md5_hex, st);
        append_record(out, relpath, seq++, last_chunk_of_file, payload + r.len0, r.len1, md5_hex, st);
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
    const std::vector<PlannedFile> &files;
    const std::vector<Batch> &batches;
    int fd;
    size_t cap;     // staging capacity of a worker
    size_t out_cap; // enough for the packed streams of any planned batch: page-locked buffers are sized once per worker
    int level;
    int device;
    std::atomic<size_t> next_batch{0};
    OrderedCommit order;
    uint64_t archive_off = 0; // guarded by the commit order
    Job(const std::string &in, const std::vector<PlannedFile> &f, const std::vector<Batch> &b, int fd_, size_t cap_, int level_, int device_)
        : input_dir(in), files(f), batches(b), fd(fd_), cap(cap_), out_cap(0), level(level_), device(device_) {
        size_t max_files = 1;
        for (const auto &x : b) max_files = std::max(max_files, x.count);
        out_cap = cap + (cap / CHUNK_SIZE + max_files + 16) * 64 + 4096;
    }
};

class Worker {
  public:
    Worker(Job &job, int id) : job_(job), ctx_(worker_ctx(job.device, id)), stage_(ctx_), out_(ctx_) {}

    void run() {
        for (;;) {
            size_t b = job_.next_batch.fetch_add(1);
            if (b >= job_.batches.size()) return;
            const Batch &batch = job_.batches[b];
            if (batch.big)
                big_file(b, job_.files[batch.first]);
            else
                small_files(b, batch);
        }
    }

  private:
    void small_files(size_t b, const Batch &batch) {
        RunStats st;
        std::ostringstream log;
        stage_.reserve(job_.cap + 64);
        // ---- read
        double t0 = now_seconds();
        std::vector<LoadedFile> loaded;
        uint64_t used = 0;
        for (size_t i = batch.first; i < batch.first + batch.count; ++i) {
            const PlannedFile &pf = job_.files[i];
            std::string full = (fs::path(job_.input_dir) / pf.relpath).string();
            std::FILE *f = std::fopen(full.c_str(), "rb");
            if (!f) { // compression.cpp:45-48: report and skip
                log << "Error opening source file: " << full << "\n";
                continue;
            }
            size_t got = pf.size ? std::fread(stage_.data() + used, 1, pf.size, f) : 0;
            std::fclose(f);
            loaded.push_back({&pf.relpath, used, got});
            used += got;
        }
        st.t_read = now_seconds() - t0;
        records_.clear();
        if (!loaded.empty()) {
            // ---- GPU
            const uint32_t nf = (uint32_t) loaded.size();
            std::vector<uint64_t> foff(nf + 1);
            for (uint32_t i = 0; i < nf; ++i) foff[i] = loaded[i].off;
            foff[nf] = used;
            const uint64_t nc = zwz_count_chunks(foff.data(), nf);
            std::vector<uint64_t> poff(nc + 1);
            std::vector<zwz_deflate_result> res(nc);
            std::vector<uint8_t> digest((size_t) nf * 16);
            out_.reserve(std::max<size_t>(job_.out_cap, used + nc * 64 + 4096));
            t0 = now_seconds();
            int rc = zwz_compress_files(ctx_, stage_.data(), foff.data(), nf, job_.level, out_.data(), out_.cap, poff.data(), res.data(),
                                        digest.data());
            if (rc != ZWZ_OK) throw std::runtime_error(std::string("zwz_compress_files: ") + zwz_last_error(ctx_));
            st.t_gpu = now_seconds() - t0;
            // ---- serialise
            t0 = now_seconds();
            size_t c = 0;
            for (uint32_t i = 0; i < nf; ++i) {
                char hex[32];
                zwz_md5_hex(&digest[(size_t) i * 16], hex);
                uint64_t nch = loaded[i].size / CHUNK_SIZE + 1;
                int seq = 0;
                for (uint64_t k = 0; k < nch; ++k, ++c)
                    emit_chunk(records_, *loaded[i].relpath, seq, k + 1 == nch, out_.data() + poff[c], res[c], hex, st);
                st.files++;
                st.raw_bytes += loaded[i].size;
                if (config().verbose) log << "md5 value size: " << MD5_DATA_SIZE << "\n"; // compression.cpp:100
            }
            st.t_write = now_seconds() - t0;
        }
        // ---- commit: take this batch's place in the archive, in batch order; the write itself runs unordered
        job_.order.wait_turn(b);
        uint64_t at = job_.archive_off;
        job_.archive_off += records_.size();
        std::string text = log.str();
        if (!text.empty()) std::cerr << text << std::flush;
        job_.order.done(b);
        t0 = now_seconds();
        write_at(job_.fd, records_.data(), records_.size(), at);
        st.t_write += now_seconds() - t0;
        merge_stats(st);
    }

    // A file larger than the staging buffer: segments of k x 65 535 bytes with k a multiple of 64, so every segment but
    // the last is also a whole number of 64-byte MD5 blocks and the digest can be chained through zwz_md5_update_device.
    // The worker keeps its turn for the whole file (its records must stay contiguous and their total size is only known
    // at the end).
    void big_file(size_t b, const PlannedFile &pf) {
        RunStats st;
        std::string full = (fs::path(job_.input_dir) / pf.relpath).string();
        job_.order.wait_turn(b);
        std::FILE *f = std::fopen(full.c_str(), "rb");
        if (!f) {
            std::cerr << "Error opening source file: " << full << std::endl;
            job_.order.done(b);
            return;
        }
        void *dev = nullptr;
        try {
            stage_.reserve(job_.cap + 64);
            const uint64_t seg_chunks = std::max<uint64_t>(64, (job_.cap / CHUNK_SIZE) / 64 * 64);
            const uint64_t seg_bytes = seg_chunks * CHUNK_SIZE;
            if (zwz_malloc_device(ctx_, seg_bytes + 64, &dev) != ZWZ_OK) throw std::runtime_error("zwz: device allocation failed");
            std::vector<uint8_t> host_seg; // staging may be smaller than a segment when batch_bytes is tiny
            uint8_t *buf = stage_.data();
            if (seg_bytes > job_.cap) {
                host_seg.resize(seg_bytes);
                buf = host_seg.data();
            }
            uint32_t state[4];
            zwz_md5_state_init(state, 1);
            uint64_t total = 0;
            int seq = 0;
            for (;;) {
                size_t got = std::fread(buf, 1, seg_bytes, f);
                total += got;
                bool final_seg = got < seg_bytes;
                uint64_t zero = 0;
                if (got && zwz_memcpy_h2d(ctx_, dev, buf, got) != ZWZ_OK) throw std::runtime_error("zwz: h2d failed");
                char hex[32];
                if (!final_seg) {
                    uint64_t len = got;
                    if (zwz_md5_update_device(ctx_, state, (const uint8_t *) dev, &zero, &len, 1, nullptr) != ZWZ_OK)
                        throw std::runtime_error(std::string("zwz: md5 update: ") + zwz_last_error(ctx_));
                } else {
                    uint64_t fullb = got & ~(uint64_t) 63, tail = got - fullb;
                    uint8_t digest[16];
                    if (fullb && zwz_md5_update_device(ctx_, state, (const uint8_t *) dev, &zero, &fullb, 1, nullptr) != ZWZ_OK)
                        throw std::runtime_error(std::string("zwz: md5 update: ") + zwz_last_error(ctx_));
                    if (zwz_md5_final_device(ctx_, state, (const uint8_t *) dev, &fullb, &tail, &total, 1, digest, nullptr) != ZWZ_OK)
                        throw std::runtime_error(std::string("zwz: md5 final: ") + zwz_last_error(ctx_));
                    zwz_md5_hex(digest, hex);
                }
                // chunks of this segment; only the final segment carries the (possibly empty) tail chunk
                uint64_t nfull = got / CHUNK_SIZE;
                uint64_t nch = final_seg ? nfull + 1 : nfull;
                std::vector<uint64_t> off(nch), poff(nch + 1);
                std::vector<uint32_t> len(nch);
                std::vector<zwz_deflate_result> res(nch);
                for (uint64_t k = 0; k < nch; ++k) {
                    off[k] = k * CHUNK_SIZE;
                    len[k] = (uint32_t) std::min<uint64_t>(CHUNK_SIZE, got - k * CHUNK_SIZE);
                }
                out_.reserve(got + nch * 64 + 4096);
                if (nch && zwz_deflate_batch(ctx_, buf, off.data(), len.data(), (uint32_t) nch, out_.data(), out_.cap, poff.data(), res.data(),
                                             job_.level) != ZWZ_OK)
                    throw std::runtime_error(std::string("zwz_deflate_batch: ") + zwz_last_error(ctx_));
                records_.clear();
                for (uint64_t k = 0; k < nch; ++k)
                    emit_chunk(records_, pf.relpath, seq, final_seg && k + 1 == nch, out_.data() + poff[k], res[k], hex, st);
                write_at(job_.fd, records_.data(), records_.size(), job_.archive_off);
                job_.archive_off += records_.size();
                if (final_seg) break;
            }
            st.files++;
            st.raw_bytes += total;
        } catch (...) {
            std::fclose(f);
            if (dev) zwz_free_device(ctx_, dev);
            throw;
        }
        std::fclose(f);
        zwz_free_device(ctx_, dev);
        job_.order.done(b);
        merge_stats(st);
    }

    Job &job_;
    zwz_ctx *ctx_;
    PinnedBuf stage_, out_;
    std::vector<char> records_;
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
    int fd = ::open(output_filename.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) {
        std::cerr << "Rank: " << world_rank << " - Error opening archive: " << output_filename << std::endl;
        return;
    }
    std::cout << "Max record line num: " << count_non_empty_lines(file_record) << std::endl;

    // the device comes up while the files are dealt and sized
    std::exception_ptr warm_error;
    std::thread warm([&] {
        try {
            ctx_for(cfg.device);
        } catch (...) {
            warm_error = std::current_exception();
        }
    });
    // the deal, then the plan: whole files per batch, in deal order
    std::vector<PlannedFile> files;
    int file_number = 0, next_file_number = world_rank;
    std::string file_path;
    while (std::getline(record_file, file_path)) {
        if (file_number == next_file_number) {
            next_file_number += cfg.world_size;
            std::error_code ec;
            uint64_t size = fs::file_size(fs::path(input_dir) / file_path, ec);
            files.push_back({file_path, ec ? 0 : size});
        }
        file_number++;
    }
    const size_t cap = cfg.batch_bytes;
    std::vector<Batch> batches;
    for (size_t i = 0; i < files.size(); ++i) {
        if (files[i].size + 64 > cap) {
            batches.push_back({i, 1, files[i].size, true});
            continue;
        }
        if (batches.empty() || batches.back().big || batches.back().bytes + files[i].size > cap || batches.back().count >= 262144)
            batches.push_back({i, 0, 0, false});
        batches.back().count++;
        batches.back().bytes += files[i].size;
    }

    warm.join();
    if (warm_error) {
        ::close(fd);
        std::rethrow_exception(warm_error);
    }
    Job job(input_dir, files, batches, fd, cap, cfg.level, cfg.device);
    const int workers = (int) std::min<size_t>((size_t) worker_count(), std::max<size_t>(1, batches.size()));
    try {
        run_workers(workers, job.order, [&](int w) {
            Worker worker(job, w);
            worker.run();
        });
    } catch (...) {
        ::close(fd);
        throw;
    }
    ::close(fd);
    std::cout << "Rank: " << world_rank << " - Total processed file: " << file_number << std::endl;
    print_timing("compress");
}

} // namespace zwzhost
