// ProjectiveICP host class (reference: src/projective_icp.cpp:66-212).  The settings live here; the loop itself is
// one cooperative kernel launch in the library — no stream helper, pinned buffer or host-side OpenCV solve.
#include <algorithm>
#include <tfusion/cuda/projective_icp.hpp>
#include "detail.hpp"

namespace tfusion {
namespace cuda {

ProjectiveICP::ProjectiveICP() : angle_thres_(deg2rad(20.f)), dist_thres_(0.1f) {
    const int iters[] = {10, 5, 4, 0};
    setIterationsNum(std::vector<int>(iters, iters + 4));
}
ProjectiveICP::~ProjectiveICP() {}

float ProjectiveICP::getDistThreshold() const { return dist_thres_; }
void ProjectiveICP::setDistThreshold(float distance) { dist_thres_ = distance; }
float ProjectiveICP::getAngleThreshold() const { return angle_thres_; }
void ProjectiveICP::setAngleThreshold(float angle) { angle_thres_ = angle; }

void ProjectiveICP::setIterationsNum(const std::vector<int>& iters) {
    iters_.assign(MAX_PYRAMID_LEVELS, 0);
    std::copy(iters.begin(), iters.begin() + std::min<size_t>(iters.size(), MAX_PYRAMID_LEVELS), iters_.begin());
}

int ProjectiveICP::getUsedLevelsNum() const {
    int i = MAX_PYRAMID_LEVELS - 1;
    for (; i >= 0 && !iters_[i]; --i) {}
    return i + 1;
}

bool ProjectiveICP::estimateTransform(Affine3f&, const Intr&, const Frame&, const Frame&) {
    error("estimateTransform(Frame, Frame) is not implemented (nor is it in the reference)", __FILE__, __LINE__);
    return false;
}

bool ProjectiveICP::estimateTransform(Affine3f&, const Intr&, const DepthPyr&, const NormalsPyr, const DepthPyr, const NormalsPyr) {
    error("the depth-pyramid ICP variant is compiled out of the reference (USE_DEPTH is undefined); use the points variant",
          __FILE__, __LINE__);
    return false;
}

bool ProjectiveICP::estimateTransform(Affine3f& affine, const Intr& intr, const PointsPyr& vcurr, const NormalsPyr ncurr,
                                      const PointsPyr vprev, const NormalsPyr nprev) {
    const int levels = getUsedLevelsNum();
    const float *vc[MAX_PYRAMID_LEVELS] = {0}, *nc[MAX_PYRAMID_LEVELS] = {0}, *vp[MAX_PYRAMID_LEVELS] = {0}, *np[MAX_PYRAMID_LEVELS] = {0};
    for (int l = 0; l < levels; ++l) {
        vc[l] = (const float*)vcurr[l].ptr(); nc[l] = (const float*)ncurr[l].ptr();
        vp[l] = (const float*)vprev[l].ptr(); np[l] = (const float*)nprev[l].ptr();
    }
    const float k[4] = {intr.fx, intr.fy, intr.cx, intr.cy};
    float aff[16];
    int ok = 0;
    TF_CHECK(tfb_icp_estimate_ext(detail::util_ctx(), levels, vc, nc, vp, np, vprev[0].cols(), vprev[0].rows(), &iters_[0], dist_thres_,
                                  angle_thres_, k, aff, &ok));
    // like the reference (projective_icp.cpp:174,197-209), `affine` holds the product of the iterations that succeeded —
    // the whole estimate when ok, Identity or a partial product when tracking failed
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) affine.matrix(r, c) = aff[r * 4 + c];
    return ok != 0;
}

}  // namespace cuda
}  // namespace tfusion
