// ThreadSanitizer harness for io::FrameRing (tests/test_frame_ring.py::test_ring_is_race_free): several decoders, a consumer
// that holds frames, destruction while the decoders are ahead.  The page-locking entry points are stubbed to fail, so the
// ring runs on pageable memory and the harness links nothing but src/frame_ring.cpp.
#include <io/frame_ring.hpp>
#include <tfusion_b200.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
extern "C" int tfb_host_alloc_pinned(void** p, size_t) { *p = nullptr; return TFB_ERR_NOMEM; }
extern "C" int tfb_host_free_pinned(void*) { return 0; }
int main(int argc, char** argv) {
    std::string dir = argc > 1 ? argv[1] : "/tmp/tsan_pgm";
    if (system(("mkdir -p " + dir).c_str())) return 2;
    const int W = 64, H = 48, N = 200;
    for (int i = 0; i < N; ++i) {
        char name[64]; snprintf(name, sizeof name, "/%04d.pgm", i);
        FILE* f = fopen((dir + name).c_str(), "wb");
        fprintf(f, "P5\n%d %d\n65535\n", W, H);
        for (int k = 0; k < W * H; ++k) { unsigned v = (unsigned)(i * 7 + k) & 0xffff; fputc(v >> 8, f); fputc(v & 255, f); }
        fclose(f);
    }
    long bad = 0;
    for (int dec = 1; dec <= 4; ++dec) {
        tfusion::io::FrameRing ring(dir, 5, 0, -1, true, dec);
        int expect = 0;
        std::vector<const tfusion::io::HostFrame*> held;
        while (const tfusion::io::HostFrame* f = ring.next()) {
            if (f->index != expect) ++bad;
            for (int k = 0; k < W * H; k += 97) if (f->data[k] != (unsigned short)((expect * 7 + k) & 0xffff)) ++bad;
            ++expect;
            held.push_back(f);
            if (held.size() > 2) { ring.release(held.front()); held.erase(held.begin()); }
        }
        for (auto* f : held) ring.release(f);
        if (expect != N) ++bad;
        printf("decoders %d: %d frames, wait %.1f ms, error '%s'\n", dec, expect, ring.consumerWaitMs(), ring.error().c_str());
    }
    { tfusion::io::FrameRing ring(dir, 3, 0, -1, true, 3); ring.next(); }   // destroyed while decoders are ahead and a frame is held
    if (system(("rm -rf " + dir).c_str())) return 2;
    printf("bad = %ld\n", bad);
    return bad != 0;
}
