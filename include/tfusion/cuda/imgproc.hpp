// Image-processing free functions (reference: include/tfusion/cuda/imgproc.hpp:9-31).  The renderers the reference
// declares here (renderImage from depth/points, renderTangentColors) are dead code on the frame path and are not
// provided; TopFu::renderImage is.
#pragma once
#include <tfusion/types.hpp>

namespace tfusion {
namespace cuda {
KF_EXPORTS void depthBilateralFilter(const Depth& in, Depth& out, int ksz, float sigma_spatial, float sigma_depth);
KF_EXPORTS void depthTruncation(Depth& depth, float threshold);
KF_EXPORTS void depthBuildPyramid(const Depth& depth, Depth& pyramid, float sigma_depth);
KF_EXPORTS void computePointNormals(const Intr& intr, const Depth& depth, Cloud& points, Normals& normals);
KF_EXPORTS void computeDists(const Depth& depth, Dists& dists, const Intr& intr);
KF_EXPORTS void resizePointsNormals(const Cloud& points, const Normals& normals, Cloud& points_out, Normals& normals_out);
KF_EXPORTS void waitAllDefaultStream();
}  // namespace cuda
}  // namespace tfusion
