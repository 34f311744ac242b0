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
        std::vector<unsigned char> raw((size_t)cols * 2);
        for (int y = 0; y < rows; ++y) {
            if (std::fread(raw.data(), 1, raw.size(), f) != raw.size()) { ok = false; break; }
            unsigned short* row = reinterpret_cast<unsigned short*>(reinterpret_cast<char*>(dst) + (size_t)y * dst_step);
            for (int x = 0; x < cols; ++x) row[x] = (unsigned short)((raw[2 * x] << 8) | raw[2 * x + 1]);   // big-endian samples
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

FrameRing::FrameRing(const std::string& dir, int slots, int first, int count, bool allow_pageable)
    : dir_(dir), first_(first), count_(count), cols_(0), rows_(0), pinned_(true), stop_(false), done_(false), produce_at_(0),
      consume_at_(0), wait_ms_(0.0) {
    if (slots < 2) slots = 2;
    if (count_ == 0 || !probePgm16(path(first_), cols_, rows_)) {
        if (count_ != 0) error_ = "cannot read " + path(first_);
        done_ = true;
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
    producer_ = std::thread(&FrameRing::produce, this);
}

FrameRing::~FrameRing() {
    {
        std::lock_guard<std::mutex> g(m_);
        stop_ = true;
    }
    cv_.notify_all();
    if (producer_.joinable()) producer_.join();
    for (Slot& s : slots_) {
        if (pinned_) tfb_host_free_pinned(s.mem);
        else std::free(s.mem);
    }
}

void FrameRing::produce() {
    const int n = (int)slots_.size();
    for (int pos = 0;; ++pos) {
        Slot& s = slots_[pos % n];
        {
            std::unique_lock<std::mutex> g(m_);
            cv_.wait(g, [&] { return stop_ || s.state == FREE; });
            if (stop_) return;
            if (count_ >= 0 && pos >= count_) { done_ = true; cv_.notify_all(); return; }
            s.state = FILLING;
        }
        const int index = first_ + pos;
        const std::string file = path(index);
        const bool ok = readPgm16(file, s.mem, s.frame.step, cols_, rows_);
        std::lock_guard<std::mutex> g(m_);
        if (!ok) {
            // an open-ended sequence ends at the first file that is not there; anything else is an error
            FILE* f = std::fopen(file.c_str(), "rb");
            if (f) std::fclose(f);
            if (f || count_ >= 0) error_ = "cannot read " + file;
            s.state = FREE;
            done_ = true;
            cv_.notify_all();
            return;
        }
        s.frame.index = index;
        s.state = READY;
        produce_at_ = pos + 1;
        cv_.notify_all();
    }
}

const HostFrame* FrameRing::next() {
    if (slots_.empty()) return nullptr;
    Slot& s = slots_[consume_at_ % (int)slots_.size()];
    std::unique_lock<std::mutex> g(m_);
    if (s.state != READY && !(done_ && produce_at_ == consume_at_)) {
        const auto t0 = std::chrono::steady_clock::now();
        cv_.wait(g, [&] { return s.state == READY || (done_ && produce_at_ == consume_at_); });
        wait_ms_ += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    if (s.state != READY) return nullptr;
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
void* tfio_ring_open(const char* dir, int slots, int first, int count, int allow_pageable) {
    if (!dir) return nullptr;
    try {
        RingHandle* h = new RingHandle();
        h->ring = new FrameRing(dir, slots, first, count, allow_pageable != 0);
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
