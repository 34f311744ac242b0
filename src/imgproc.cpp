// Stand-alone image stages (reference: src/imgproc.cpp:3-77 -> src/cuda/imgproc.cu).  Each wrapper sizes its output
// like the reference and forwards to the C ABI; images must be dense (DeviceArray2D here always is).
#include <tfusion/cuda/imgproc.hpp>
#include "detail.hpp"

namespace tfusion {
namespace cuda {

static void need_dense(const DeviceMemory2D& m, const char* who) {
    if (!m.empty() && m.step() != (size_t)m.colsBytes()) error("pitched images are not supported by the stand-alone stages", __FILE__, __LINE__, who);
}

void depthBilateralFilter(const Depth& in, Depth& out, int kernel_size, float sigma_spatial, float sigma_depth) {
    out.create(in.rows(), in.cols());
    need_dense(in, "depthBilateralFilter");
    TF_CHECK(tfb_bilateral_filter(detail::util_ctx(), in.ptr(), out.ptr(), in.cols(), in.rows(), kernel_size, sigma_spatial, sigma_depth));
}

void depthTruncation(Depth& depth, float threshold) {
    need_dense(depth, "depthTruncation");
    TF_CHECK(tfb_truncate_depth(detail::util_ctx(), depth.ptr(), depth.cols(), depth.rows(), threshold));
}

void depthBuildPyramid(const Depth& depth, Depth& pyramid, float sigma_depth) {
    pyramid.create(depth.rows() / 2, depth.cols() / 2);
    need_dense(depth, "depthBuildPyramid");
    TF_CHECK(tfb_depth_pyr(detail::util_ctx(), depth.ptr(), pyramid.ptr(), depth.cols(), depth.rows(), sigma_depth));
}

void computePointNormals(const Intr& intr, const Depth& depth, Cloud& points, Normals& normals) {
    points.create(depth.rows(), depth.cols());
    normals.create(depth.rows(), depth.cols());
    need_dense(depth, "computePointNormals");
    TF_CHECK(tfb_compute_point_normals(detail::util_ctx(), depth.ptr(), (float*)points.ptr(), (float*)normals.ptr(), depth.cols(),
                                       depth.rows(), intr.fx, intr.fy, intr.cx, intr.cy));
}

void computeDists(const Depth& depth, Dists& dists, const Intr&) {
    // the reference computes a per-pixel ray length factor from the intrinsics and then ignores it (imgproc.cu:272-277)
    dists.create(depth.rows(), depth.cols());
    need_dense(depth, "computeDists");
    TF_CHECK(tfb_compute_dists(detail::util_ctx(), depth.ptr(), dists.ptr(), depth.cols(), depth.rows()));
}

void resizePointsNormals(const Cloud& points, const Normals& normals, Cloud& points_out, Normals& normals_out) {
    points_out.create(points.rows() / 2, points.cols() / 2);
    normals_out.create(normals.rows() / 2, normals.cols() / 2);
    TF_CHECK(tfb_resize_points_normals(detail::util_ctx(), (const float*)points.ptr(), (const float*)normals.ptr(),
                                       (float*)points_out.ptr(), (float*)normals_out.ptr(), points.cols(), points.rows()));
}

void waitAllDefaultStream() { TF_CHECK(tfb_sync(detail::util_ctx())); }

}  // namespace cuda
}  // namespace tfusion
