// compress_pipeline.cpp — the compress side of one rank (compression.cpp:24-194 in the reference).
//
// The reference runs a producer thread that cuts files into 65 535-byte chunks, a mutex queue, ONE consumer thread that
// calls zlib per chunk, and a writer that re-reads each source file for its MD5. Here the rank's files are read into a
// pinned staging buffer batch by batch and each batch makes ONE trip through the GPU (zwz_compress_files: upload once,
// deflate every chunk, MD5 every file from the same resident bytes, pack, download); the host then serialises records in
// the reference's byte layout (compression.cpp:73-104). No queue, no lock: record order is file order then sequence order,
// which is exactly what the reference produces with NUM_CONSUMERS = 1.
#include "zwz_host.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>

namespace zwzhost {

namespace fs = std::filesystem;
zwz_ctx *ctx_for(int device);

namespace {

struct PendingFile {
    std::string relpath;
    uint64_t off = 0, size = 0; // inside the staging buffer
};

// compression.cpp:73-104: i32 total_size, i32 path_len, path, i32 sequence_id, u8 is_last_chunk, payload[, 32 hex chars]
void append_record(std::vector<char> &out, const std::string &relpath, int sequence_id, bool is_last_chunk, const uint8_t *payload,
                   uint32_t payload_len, const char *md5_hex) {
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
    stats().records++;
    stats().payload_bytes += payload_len;
}

std::string generate_output_filename(const std::string &output_dir, int world_rank) { // compression.cpp:151-159
    fs::path dir(output_dir);
    dir = dir / "";
    return (dir / ("compressed_" + std::to_string(world_rank) + ".zwz")).string();
}

class Compressor {
  public:
    Compressor(int device, int level, size_t batch_bytes, std::FILE *dest) : ctx_(ctx_for(device)), level_(level), cap_(batch_bytes), dest_(dest) {
        if (zwz_malloc_pinned(ctx_, cap_ + 64, (void **) &stage_) != ZWZ_OK) throw std::runtime_error("zwz: pinned staging allocation failed");
        out_cap_ = cap_ + (cap_ / CHUNK_SIZE + 4096) * 64 + 4096;
        if (zwz_malloc_pinned(ctx_, out_cap_, (void **) &out_) != ZWZ_OK) throw std::runtime_error("zwz: pinned output allocation failed");
    }
    ~Compressor() {
        zwz_free_pinned(ctx_, stage_);
        zwz_free_pinned(ctx_, out_);
    }

    // whole files that fit the staging buffer are batched; larger ones stream through in segments
    void add_file(const std::string &relpath, const fs::path &full_path) {
        std::FILE *f = std::fopen(full_path.c_str(), "rb");
        if (!f) { // compression.cpp:45-48: report and skip
            std::cerr << "Error opening source file: " << full_path << std::endl;
            return;
        }
        std::error_code ec;
        uint64_t size = fs::file_size(full_path, ec);
        if (ec) size = 0;
        if (size + 64 > cap_) {
            flush();
            big_file(relpath, f, full_path.string());
            std::fclose(f);
            return;
        }
        if (used_ + size > cap_ || files_.size() >= 262144) flush();
        double t0 = now_seconds();
        size_t got = size ? std::fread(stage_ + used_, 1, size, f) : 0;
        std::fclose(f);
        stats().t_read += now_seconds() - t0;
        files_.push_back({relpath, used_, got});
        used_ += got;
    }

    void flush() {
        if (files_.empty()) return;
        const uint32_t nf = (uint32_t) files_.size();
        std::vector<uint64_t> foff(nf + 1);
        for (uint32_t i = 0; i < nf; ++i) foff[i] = files_[i].off;
        foff[nf] = used_;
        const uint64_t nc = zwz_count_chunks(foff.data(), nf);
        std::vector<uint64_t> poff(nc + 1);
        std::vector<zwz_deflate_result> res(nc);
        std::vector<uint8_t> digest((size_t) nf * 16);
        ensure_out(used_ + nc * 64 + 4096);
        double t0 = now_seconds();
        int rc = zwz_compress_files(ctx_, stage_, foff.data(), nf, level_, out_, out_cap_, poff.data(), res.data(), digest.data());
        if (rc != ZWZ_OK) throw std::runtime_error(std::string("zwz_compress_files: ") + zwz_last_error(ctx_));
        stats().t_gpu += now_seconds() - t0;
        t0 = now_seconds();
        records_.clear();
        size_t c = 0;
        for (uint32_t i = 0; i < nf; ++i) {
            char hex[32];
            zwz_md5_hex(&digest[(size_t) i * 16], hex);
            uint64_t nch = files_[i].size / CHUNK_SIZE + 1;
            int seq = 0;
            for (uint64_t k = 0; k < nch; ++k, ++c) emit_chunk(files_[i].relpath, seq, k + 1 == nch, out_ + poff[c], res[c], hex);
            stats().files++;
            stats().raw_bytes += files_[i].size;
            if (config().verbose) std::cout << "md5 value size: " << MD5_DATA_SIZE << std::endl; // compression.cpp:100
        }
        write_out();
        stats().t_write += now_seconds() - t0;
        files_.clear();
        used_ = 0;
    }

  private:
    // one chunk -> one record, or two when the split rule fired (the reference would have silently truncated this chunk)
    void emit_chunk(const std::string &relpath, int &seq, bool last_chunk_of_file, const uint8_t *payload, const zwz_deflate_result &r,
                    const char *md5_hex) {
        if (r.len1 == 0) {
            append_record(records_, relpath, seq++, last_chunk_of_file, payload, r.len0, md5_hex);
        } else {
            append_record(records_, relpath, seq++, false, payload, r.len0, md5_hex);
            append_record(records_, relpath, seq++, last_chunk_of_file, payload + r.len0, r.len1, md5_hex);
        }
    }
    void write_out() {
        if (!records_.empty() && std::fwrite(records_.data(), 1, records_.size(), dest_) != records_.size())
            throw std::runtime_error("zwz: short write to archive");
        records_.clear();
    }
    void ensure_out(uint64_t need) {
        if (need <= out_cap_) return;
        zwz_free_pinned(ctx_, out_);
        out_cap_ = need + need / 8;
        if (zwz_malloc_pinned(ctx_, out_cap_, (void **) &out_) != ZWZ_OK) throw std::runtime_error("zwz: pinned output allocation failed");
    }

    // A file larger than the staging buffer: segments of k x 65 535 bytes with k a multiple of 64, so every segment but
    // the last is also a whole number of 64-byte MD5 blocks and the digest can be chained through zwz_md5_update_device.
    void big_file(const std::string &relpath, std::FILE *f, const std::string &full_path) {
        const uint64_t seg_chunks = std::max<uint64_t>(64, (cap_ / CHUNK_SIZE) / 64 * 64);
        const uint64_t seg_bytes = seg_chunks * CHUNK_SIZE;
        void *dev = nullptr;
        if (zwz_malloc_device(ctx_, seg_bytes + 64, &dev) != ZWZ_OK) throw std::runtime_error("zwz: device allocation failed");
        std::vector<uint8_t> host_seg; // staging may be smaller than a segment when batch_bytes is tiny
        uint8_t *buf = stage_;
        if (seg_bytes > cap_) {
            host_seg.resize(seg_bytes);
            buf = host_seg.data();
        }
        uint32_t state[4];
        zwz_md5_state_init(state, 1);
        uint64_t total = 0;
        int seq = 0;
        // records of this file are buffered until the digest is known only for the LAST record; earlier ones stream out
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
                uint64_t full = got & ~(uint64_t) 63, tail = got - full;
                uint8_t digest[16];
                if (full && zwz_md5_update_device(ctx_, state, (const uint8_t *) dev, &zero, &full, 1, nullptr) != ZWZ_OK)
                    throw std::runtime_error(std::string("zwz: md5 update: ") + zwz_last_error(ctx_));
                if (zwz_md5_final_device(ctx_, state, (const uint8_t *) dev, &full, &tail, &total, 1, digest, nullptr) != ZWZ_OK)
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
            ensure_out(got + nch * 64 + 4096);
            if (nch && zwz_deflate_batch(ctx_, buf, off.data(), len.data(), (uint32_t) nch, out_, out_cap_, poff.data(), res.data(), level_) != ZWZ_OK)
                throw std::runtime_error(std::string("zwz_deflate_batch: ") + zwz_last_error(ctx_));
            for (uint64_t k = 0; k < nch; ++k) emit_chunk(relpath, seq, final_seg && k + 1 == nch, out_ + poff[k], res[k], hex);
            write_out();
            if (final_seg) break;
        }
        zwz_free_device(ctx_, dev);
        stats().files++;
        stats().raw_bytes += total;
        (void) full_path;
    }

    zwz_ctx *ctx_;
    int level_;
    size_t cap_;
    std::FILE *dest_;
    uint8_t *stage_ = nullptr, *out_ = nullptr;
    uint64_t out_cap_ = 0, used_ = 0;
    std::vector<PendingFile> files_;
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
    std::FILE *dest = std::fopen(output_filename.c_str(), "wb");
    if (!dest) {
        std::cerr << "Rank: " << world_rank << " - Error opening archive: " << output_filename << std::endl;
        return;
    }
    std::cout << "Max record line num: " << count_non_empty_lines(file_record) << std::endl;
    {
        Compressor comp(cfg.device, cfg.level, cfg.batch_bytes, dest);
        int file_number = 0, next_file_number = world_rank;
        std::string file_path;
        while (std::getline(record_file, file_path)) {
            if (file_number == next_file_number) {
                next_file_number += cfg.world_size;
                comp.add_file(file_path, fs::path(input_dir) / file_path);
            }
            file_number++;
        }
        comp.flush();
        std::cout << "Rank: " << world_rank << " - Total processed file: " << file_number << std::endl;
    }
    std::fclose(dest);
    print_timing("compress");
}

} // namespace zwzhost
