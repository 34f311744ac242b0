// Decode-ahead ring of page-locked depth frames (include/io/frame_ring.hpp).  Replaces, for file sequences, the
// synchronous cv::imread + DeviceArray2D::upload pair of apps/demo.cpp:91-100.
#include <io/frame_ring.hpp>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <tfusion_b200.h>

namespace tfusion {
namespace io {
namespace {

// next header token (comments run from '#' to the end of the line); returns false at end of file
bool header_int(FILE* f, int& value) {
    int c = std::fgetc(f);
    for (;;) {
        while (c == ' ' || c == '\t' || c == '\r' || c == '\n') c = std::fgetc(f);
        if (c != '#') break;
        while (c != '\n' && c != EOF) c = std::fgetc(f);
    }
    if (c < '0' || c > '9') return false;
    long v = 0;
    while (c >= '0' && c <= '9') {
        v = v * 10 + (c - '0');
        if (v > 0x7fffffffL) return false;
        c = std::fgetc(f);
    }
    std::ungetc(c, f);   // the single whitespace that ends the header is consumed by the caller
    value = (int)v;
    return true;
}

struct Header { bool plain; int cols, rows, maxval; };

// big-endian file samples -> host order (a no-op on a big-endian host)
inline void swap_rows(unsigned short* row, int n) {
#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_BIG_ENDIAN__
    (void)row; (void)n;
#else
    for (int x = 0; x < n; ++x) row[x] = __builtin_bswap16(row[x]);
#endif
}

bool read_header(FILE* f, Header& h) {
    const int p = std::fgetc(f), k = std::fgetc(f);
    if (p != 'P' || (k != '5' && k != '2')) return false;
    h.plain = (k == '2');
    if (!header_int(f, h.cols) || !header_int(f, h.rows) || !header_int(f, h.maxval)) return false;
    if (h.cols <= 0 || h.rows <= 0 || h.maxval <= 255 || h.maxval > 65535) return false;   // 16-bit depth only
    const int ws = std::fgetc(f);   // exactly one whitespace character, then the raster
    return ws == ' ' || ws == '\t' || ws == '\r' || ws == '\n';
}

}  // namespace

bool probePgm16(const std::string& path, int& cols, int& rows) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    Header h;
    const bool ok = read_header(f, h);
    std::fclose(f);
    if (ok) { cols = h.cols; rows = h.rows; }
    return ok;
}

bool readPgm16(const std::string& path, unsigned short* dst, size_t dst_step, int cols, int rows) {
    if (!dst || dst_step < (size_t)cols * 2) return false;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    Header h;
    bool ok = read_header(f, h) && h.cols == cols && h.rows == rows;
    if (ok && h.plain) {
        for (int y = 0; ok && y < rows; ++y) {
            unsigned short* row = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(dst) + (size_t)y * dst_step);
            for (int x = 0; x < cols; ++x) {
                int v;
                if (!header_int(f, v) || v > 65535) { ok = false; break; }
                row[x] = (unsigned short)v;
            }
        }
    } else if (ok) {
        // the raster goes straight into the destination (one read when the rows are dense), then the big-endian samples are
        // swapped in place — a loop the compiler turns into byte shuffles; at 4 000 frames/s the decode must not cost more
        // than the frame
        const size_t row_bytes = (size_t)cols * 2;
        if (dst_step == row_bytes) {
            ok = std::fread(dst, 1, row_bytes * rows, f) == row_bytes * rows;
        } else {
            for (int y = 0; ok && y < rows; ++y)
                ok = std::fread(reinterpret_cast<char*>(dst) + (size_t)y * dst_step, 1, row_bytes, f) == row_bytes;
        }
        for (int y = 0; ok && y < rows; ++y) {
            unsigned short* row = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(dst) + (size_t)y * dst_step);
            swap_rows(row, cols);
        }
    }
    std::fclose(f);
    return ok;
}

std::string FrameRing::path(int index) const {
    char name[32];
    std::snprintf(name, sizeof(name), "/%04d.pgm", index);   // apps/demo.cpp:93
    return dir_ + name;
}

FrameRing::FrameRing(const std::string& dir, int slots, int first, int count, bool allow_pageable, int decoders)
    : dir_(dir), first_(first), count_(count), cols_(0), rows_(0), pinned_(true), stop_(false), claim_at_(0),
      end_at_(count >= 0 ? count : 0x7fffffff), consume_at_(0), wait_ms_(0.0) {
    if (slots < 2) slots = 2;
    if (decoders < 1) decoders = 1;
    if (decoders > slots) decoders = slots;
    if (count_ == 0 || !probePgm16(path(first_), cols_, rows_)) {
        if (count_ != 0) error_ = "cannot read " + path(first_);
        end_at_ = 0;
        return;
    }
    const size_t bytes = (size_t)cols_ * rows_ * sizeof(unsigned short);
    for (int i = 0; i < slots; ++i) {
        void* p = nullptr;
        if (pinned_ && tfb_host_alloc_pinned(&p, bytes) != TFB_OK) {
            if (!allow_pageable) {
                for (Slot& s : slots_) tfb_host_free_pinned(s.mem);
                throw std::runtime_error("FrameRing: cannot page-lock host memory (no CUDA device?)");
            }
            for (Slot& s : slots_) {   // all slots of one kind
                tfb_host_free_pinned(s.mem);
                s.mem = static_cast<unsigned short*>(std::malloc(bytes));
                if (!s.mem) throw std::runtime_error("FrameRing: out of host memory");
            }
            pinned_ = false;
        }
        if (!pinned_) p = std::malloc(bytes);
        if (!p) throw std::runtime_error("FrameRing: out of host memory");
        Slot s;
        s.mem = static_cast<unsigned short*>(p);
        s.frame = HostFrame{s.mem, rows_, cols_, (size_t)cols_ * sizeof(unsigned short), -1};
        s.state = FREE;
        slots_.push_back(s);
    }
    for (Slot& s : slots_) s.frame.data = s.mem;
    for (int i = 0; i < decoders; ++i) decoders_.emplace_back(&FrameRing::produce, this);
}

FrameRing::~FrameRing() {
    {
        std::lock_guard<std::mutex> g(m_);
        stop_ = true;
    }
    cv_.notify_all();
    for (std::thread& t : decoders_)
        if (t.joinable()) t.join();
    for (Slot& s : slots_) {
        if (pinned_) tfb_host_free_pinned(s.mem);
        else std::free(s.mem);
    }
}

// A decoder thread: claim the next position of the sequence once its slot is free, decode, mark the slot ready.  Positions
// are claimed in order and a slot serves positions p, p + slots, ...: the consumer frees them in order, so with several
// decoders the files are still delivered in sequence.  The first position that cannot be read ends the sequence there
// (end_at_), whatever later positions other decoders have already finished.
void FrameRing::produce() {
    const int n = (int)slots_.size();
    for (;;) {
        int pos;
        {
            std::unique_lock<std::mutex> g(m_);
            cv_.wait(g, [&] { return stop_ || claim_at_ >= end_at_ || slots_[claim_at_ % n].state == FREE; });
            if (stop_ || claim_at_ >= end_at_) return;
            pos = claim_at_++;
            slots_[pos % n].state = FILLING;
        }
        Slot& s = slots_[pos % n];
        const std::string file = path(first_ + pos);
        const bool ok = readPgm16(file, s.mem, s.frame.step, cols_, rows_);
        std::lock_guard<std::mutex> g(m_);
        if (!ok) {
            if (pos < end_at_) {
                end_at_ = pos;
                // an open-ended sequence ends at the first file that is not there; anything else is an error
                FILE* f = std::fopen(file.c_str(), "rb");
                if (f) std::fclose(f);
                error_ = (f || count_ >= 0) ? "cannot read " + file : std::string();
            }
            s.state = FREE;
            cv_.notify_all();
            return;
        }
        s.frame.index = first_ + pos;
        s.state = READY;
        cv_.notify_all();
    }
}

const HostFrame* FrameRing::next() {
    if (slots_.empty()) return nullptr;
    Slot& s = slots_[consume_at_ % (int)slots_.size()];
    std::unique_lock<std::mutex> g(m_);
    auto settled = [&] { return consume_at_ >= end_at_ || (s.state == READY && s.frame.index == first_ + consume_at_); };
    if (!settled()) {
        const auto t0 = std::chrono::steady_clock::now();
        cv_.wait(g, settled);
        wait_ms_ += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    if (consume_at_ >= end_at_) return nullptr;
    s.state = HELD;
    ++consume_at_;
    return &s.frame;
}

void FrameRing::release(const HostFrame* frame) {
    if (!frame) return;
    {
        std::lock_guard<std::mutex> g(m_);
        for (Slot& s : slots_)
            if (&s.frame == frame && s.state == HELD) s.state = FREE;
    }
    cv_.notify_all();
}

}  // namespace io
}  // namespace tfusion

using tfusion::io::FrameRing;
using tfusion::io::HostFrame;

namespace {
struct RingHandle {
    FrameRing* ring;
    std::vector<const HostFrame*> held;
};
}  // namespace

extern "C" {

int tfio_probe_pgm16(const char* path, int* cols, int* rows) {
    return (path && cols && rows && tfusion::io::probePgm16(path, *cols, *rows)) ? 1 : 0;
}
int tfio_read_pgm16(const char* path, unsigned short* dst, size_t dst_step, int cols, int rows) {
    return (path && tfusion::io::readPgm16(path, dst, dst_step, cols, rows)) ? 1 : 0;
}
void* tfio_ring_open(const char* dir, int slots, int first, int count, int allow_pageable, int decoders) {
    if (!dir) return nullptr;
    try {
        RingHandle* h = new RingHandle();
        h->ring = new FrameRing(dir, slots, first, count, allow_pageable != 0, decoders);
        return h;
    } catch (const std::exception&) {
        return nullptr;
    }
}
int tfio_ring_next(void* ring, const unsigned short** data, int* rows, int* cols, size_t* step, int* index) {
    RingHandle* h = static_cast<RingHandle*>(ring);
    if (!h || !data || !rows || !cols || !step || !index) return 0;
    const HostFrame* f = h->ring->next();
    if (!f) return 0;
    h->held.push_back(f);
    *data = f->data; *rows = f->rows; *cols = f->cols; *step = f->step; *index = f->index;
    return 1;
}
void tfio_ring_release(void* ring, int index) {
    RingHandle* h = static_cast<RingHandle*>(ring);
    if (!h) return;
    for (size_t i = 0; i < h->held.size(); ++i)
        if (h->held[i]->index == index) {
            h->ring->release(h->held[i]);
            h->held.erase(h->held.begin() + (long)i);
            return;
        }
}
int tfio_ring_pinned(void* ring) { return ring && static_cast<RingHandle*>(ring)->ring->pinned() ? 1 : 0; }
const char* tfio_ring_error(void* ring) { return ring ? static_cast<RingHandle*>(ring)->ring->error().c_str() : "null ring"; }
void tfio_ring_close(void* ring) {
    RingHandle* h = static_cast<RingHandle*>(ring);
    if (!h) return;
    delete h->ring;
    delete h;
}

}  // extern "C"
