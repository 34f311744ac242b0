// Input side of the path: what apps/demo.cpp does per frame in front of TopFu::operator() —
//     depth = cv::imread("%04d.pgm", CV_16U); depth_device_.upload(depth.data, depth.step, depth.rows, depth.cols);
// (apps/demo.cpp:91-100) — as a decode-ahead ring of page-locked frames.  Decoder threads read the next files while the
// GPU works on the current one; the consumer hands a slot straight to TopFu::operator()(const io::HostFrame&), whose
// upload is an asynchronous copy on the library's second stream, beside the previous frame's integration and raycast.
// At a few thousand frames per second the file decode and the 0.6 MB copy are the next bottleneck; the reference does both
// synchronously on the frame's critical path.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <tfusion/exports.hpp>

namespace tfusion {
namespace io {

// a 16-bit depth frame in host memory (millimetres, 0 = no return), rows x cols, `step` bytes per row
struct HostFrame {
    const unsigned short* data;
    int rows, cols;
    size_t step;
    int index;   // position in the sequence
};

// Netpbm greyscale reader with cv::imread(path, CV_16U)'s result for a depth file: P5 (binary) and P2 (plain), `#` comments
// in the header, samples big-endian, NO rescaling by maxval (OpenCV does none).  Only 16-bit files (maxval > 255) are
// accepted: OpenCV would hand back an 8-bit image for the others, which the reference then uploads as if it were 16-bit.
// probe: header only.  read: into dst (dst_step bytes per row, >= cols * 2).  Both return false on any malformed input.
KF_EXPORTS bool probePgm16(const std::string& path, int& cols, int& rows);
KF_EXPORTS bool readPgm16(const std::string& path, unsigned short* dst, size_t dst_step, int cols, int rows);

// Frames <dir>/%04d.pgm, first .. first + count - 1 (count < 0: until the first missing file), decoded ahead of the consumer
// into `slots` page-locked buffers (cudaHostAlloc through the C ABI: the copy engine reads them without a staging copy) by
// `decoders` threads.  Measured on this image's host, 640x480 files on a RAM disk, consumer only releasing: 4 500 frames/s
// with one decoder (224 us per file: open, one 614 KB read, byte swap), 8 900 with two, 18 000 with four.  One decoder is
// about the GPU's frame rate, hence two by default.
class KF_EXPORTS FrameRing {
public:
    // allow_pageable: fall back to ordinary memory when page-locking is impossible (no CUDA device: the CPU test suite).
    // The default fails loudly instead, like everything else on the product path.
    FrameRing(const std::string& dir, int slots = 4, int first = 0, int count = -1, bool allow_pageable = false, int decoders = 2);
    ~FrameRing();

    // the next frame of the sequence, in order; blocks until it is decoded.  nullptr at the end of the sequence (or after a
    // file that could not be read: see error()).  The frame stays valid until release().
    const HostFrame* next();
    void release(const HostFrame* frame);

    bool pinned() const { return pinned_; }
    int cols() const { return cols_; }
    int rows() const { return rows_; }
    const std::string& error() const { return error_; }
    // how long the consumer waited for the decoder in next(), in total (0 when the ring always ran ahead)
    double consumerWaitMs() const { return wait_ms_; }

private:
    FrameRing(const FrameRing&);
    FrameRing& operator=(const FrameRing&);
    void produce();
    std::string path(int index) const;

    enum State { FREE, FILLING, READY, HELD };
    struct Slot {
        unsigned short* mem;
        HostFrame frame;
        State state;
    };
    std::string dir_, error_;
    int first_, count_, cols_, rows_;
    bool pinned_, stop_;
    std::vector<Slot> slots_;
    // sequence positions (slot = position % slots): next to be claimed by a decoder, first that will never be delivered
    // (the count, or the first unreadable file), next to be handed to the consumer
    int claim_at_, end_at_, consume_at_;
    double wait_ms_;
    std::mutex m_;
    std::condition_variable cv_;
    std::vector<std::thread> decoders_;
};

}  // namespace io
}  // namespace tfusion

// The same through a C interface (ctypes-friendly; tests/test_frame_ring.py)
extern "C" {
KF_EXPORTS int tfio_probe_pgm16(const char* path, int* cols, int* rows);
KF_EXPORTS int tfio_read_pgm16(const char* path, unsigned short* dst, size_t dst_step, int cols, int rows);
KF_EXPORTS void* tfio_ring_open(const char* dir, int slots, int first, int count, int allow_pageable, int decoders);
KF_EXPORTS int tfio_ring_next(void* ring, const unsigned short** data, int* rows, int* cols, size_t* step, int* index);   // 1 frame, 0 end
KF_EXPORTS void tfio_ring_release(void* ring, int index);
KF_EXPORTS int tfio_ring_pinned(void* ring);
KF_EXPORTS const char* tfio_ring_error(void* ring);
KF_EXPORTS void tfio_ring_close(void* ring);
}
