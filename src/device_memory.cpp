// DeviceMemory / DeviceMemory2D over the C ABI (reference: src/device_memory.cpp — same ownership rules: copies
// share, the last owner frees, user-provided memory is never freed).
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <utility>
#include <tfusion/cuda/device_memory.hpp>
#include "detail.hpp"

namespace tfusion {

void cuda::error(const char* error_string, const char* file, const int line, const char* func) {
    // the reference prints "KinFu2 error" and exits (src/device_memory.cpp:7-11); keep that contract for the C++ API
    std::fprintf(stderr, "tfusion error: %s\t%s:%d %s\n", error_string, file, line, func ? func : "");
    std::exit(EXIT_FAILURE);
}

namespace detail {
tfb_ctx* util_ctx() {
    // created once, by whichever thread gets here first; lives as long as the process (it owns the stream that stand-alone
    // DeviceMemory / imgproc / ProjectiveICP calls run on, and static destructors may still use it)
    static tfb_ctx* c = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        tfb_params p;
        tfb_default_params(&p);
        p.cols = 8; p.rows = 8; p.num_blocks = 1; p.num_buckets = 2; p.excess_size = 1;
        int rc = tfb_create(&p, nullptr, &c);
        if (rc != TFB_OK) cuda::error("cannot create a CUDA context (no device? there is no CPU fallback)", __FILE__, __LINE__);
    });
    return c;
}
void check(int rc, const char* what, const char* file, int line) {
    if (rc != TFB_OK) cuda::error(what, file, line);
}
}  // namespace detail

namespace cuda {

DeviceMemory::DeviceMemory() : data_(0), sizeBytes_(0), refcount_(0) {}
DeviceMemory::DeviceMemory(void* ptr_arg, size_t sizeBytes_arg) : data_(ptr_arg), sizeBytes_(sizeBytes_arg), refcount_(0) {}
DeviceMemory::DeviceMemory(size_t sizeBytes_arg) : data_(0), sizeBytes_(0), refcount_(0) { create(sizeBytes_arg); }
DeviceMemory::~DeviceMemory() { release(); }
DeviceMemory::DeviceMemory(const DeviceMemory& o) : data_(o.data_), sizeBytes_(o.sizeBytes_), refcount_(o.refcount_) {
    if (refcount_) __atomic_add_fetch(refcount_, 1, __ATOMIC_RELAXED);
}
DeviceMemory& DeviceMemory::operator=(const DeviceMemory& o) {
    if (this != &o) {
        if (o.refcount_) __atomic_add_fetch(o.refcount_, 1, __ATOMIC_RELAXED);
        release();
        data_ = o.data_; sizeBytes_ = o.sizeBytes_; refcount_ = o.refcount_;
    }
    return *this;
}
void DeviceMemory::create(size_t sizeBytes_arg) {
    if (sizeBytes_arg == sizeBytes_) return;
    if (sizeBytes_arg > 0) {
        release();
        sizeBytes_ = sizeBytes_arg;
        TF_CHECK(tfb_dev_alloc(&data_, sizeBytes_));
        refcount_ = new int(1);
    }
}
void DeviceMemory::release() {
    if (refcount_ && __atomic_sub_fetch(refcount_, 1, __ATOMIC_ACQ_REL) == 0) {
        delete refcount_;
        tfb_dev_free(data_);
    }
    data_ = 0; sizeBytes_ = 0; refcount_ = 0;
}
void DeviceMemory::copyTo(DeviceMemory& other) const {
    if (empty()) { other.release(); return; }
    other.create(sizeBytes_);
    TF_CHECK(tfb_memcpy_d2d(detail::util_ctx(), other.data_, data_, sizeBytes_));
    TF_CHECK(tfb_sync(detail::util_ctx()));
}
void DeviceMemory::upload(const void* host_ptr_arg, size_t sizeBytes_arg) {
    create(sizeBytes_arg);
    TF_CHECK(tfb_h2d(detail::util_ctx(), data_, host_ptr_arg, sizeBytes_));
    TF_CHECK(tfb_sync(detail::util_ctx()));
}
void DeviceMemory::download(void* host_ptr_arg) const { TF_CHECK(tfb_d2h(detail::util_ctx(), host_ptr_arg, data_, sizeBytes_)); }
void DeviceMemory::swap(DeviceMemory& o) { std::swap(data_, o.data_); std::swap(sizeBytes_, o.sizeBytes_); std::swap(refcount_, o.refcount_); }
bool DeviceMemory::empty() const { return !data_; }
size_t DeviceMemory::sizeBytes() const { return sizeBytes_; }

DeviceMemory2D::DeviceMemory2D() : data_(0), step_(0), colsBytes_(0), rows_(0), refcount_(0) {}
DeviceMemory2D::DeviceMemory2D(int rows_arg, int colsBytes_arg) : data_(0), step_(0), colsBytes_(0), rows_(0), refcount_(0) {
    create(rows_arg, colsBytes_arg);
}
DeviceMemory2D::DeviceMemory2D(int rows_arg, int colsBytes_arg, void* data_arg, size_t step_arg)
    : data_(data_arg), step_(step_arg), colsBytes_(colsBytes_arg), rows_(rows_arg), refcount_(0) {}
DeviceMemory2D::~DeviceMemory2D() { release(); }
DeviceMemory2D::DeviceMemory2D(const DeviceMemory2D& o)
    : data_(o.data_), step_(o.step_), colsBytes_(o.colsBytes_), rows_(o.rows_), refcount_(o.refcount_) {
    if (refcount_) __atomic_add_fetch(refcount_, 1, __ATOMIC_RELAXED);
}
DeviceMemory2D& DeviceMemory2D::operator=(const DeviceMemory2D& o) {
    if (this != &o) {
        if (o.refcount_) __atomic_add_fetch(o.refcount_, 1, __ATOMIC_RELAXED);
        release();
        data_ = o.data_; step_ = o.step_; colsBytes_ = o.colsBytes_; rows_ = o.rows_; refcount_ = o.refcount_;
    }
    return *this;
}
void DeviceMemory2D::create(int rows_arg, int colsBytes_arg) {
    if (colsBytes_ == colsBytes_arg && rows_ == rows_arg) return;
    if (rows_arg > 0 && colsBytes_arg > 0) {
        release();
        colsBytes_ = colsBytes_arg; rows_ = rows_arg;
        step_ = (size_t)colsBytes_;  // dense on purpose, see the header
        TF_CHECK(tfb_dev_alloc(&data_, step_ * (size_t)rows_));
        refcount_ = new int(1);
    }
}
void DeviceMemory2D::release() {
    if (refcount_ && __atomic_sub_fetch(refcount_, 1, __ATOMIC_ACQ_REL) == 0) {
        delete refcount_;
        tfb_dev_free(data_);
    }
    data_ = 0; step_ = 0; colsBytes_ = 0; rows_ = 0; refcount_ = 0;
}
void DeviceMemory2D::copyTo(DeviceMemory2D& other) const {
    if (empty()) { other.release(); return; }
    other.create(rows_, colsBytes_);
    TF_CHECK(tfb_memcpy_2d(detail::util_ctx(), other.data_, other.step_, data_, step_, (size_t)colsBytes_, rows_, 2));
    TF_CHECK(tfb_sync(detail::util_ctx()));
}
void DeviceMemory2D::upload(const void* host_ptr_arg, size_t host_step_arg, int rows_arg, int colsBytes_arg) {
    create(rows_arg, colsBytes_arg);
    TF_CHECK(tfb_memcpy_2d(detail::util_ctx(), data_, step_, host_ptr_arg, host_step_arg, (size_t)colsBytes_, rows_, 0));
}
void DeviceMemory2D::download(void* host_ptr_arg, size_t host_step_arg) const {
    TF_CHECK(tfb_memcpy_2d(detail::util_ctx(), host_ptr_arg, host_step_arg, data_, step_, (size_t)colsBytes_, rows_, 1));
}
void DeviceMemory2D::swap(DeviceMemory2D& o) {
    std::swap(data_, o.data_); std::swap(step_, o.step_); std::swap(colsBytes_, o.colsBytes_); std::swap(rows_, o.rows_);
    std::swap(refcount_, o.refcount_);
}
bool DeviceMemory2D::empty() const { return !data_; }
int DeviceMemory2D::colsBytes() const { return colsBytes_; }
int DeviceMemory2D::rows() const { return rows_; }
size_t DeviceMemory2D::step() const { return step_; }

}  // namespace cuda
}  // namespace tfusion
