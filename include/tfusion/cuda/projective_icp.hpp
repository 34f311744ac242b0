// Projective point-to-plane ICP front end (reference: include/tfusion/cuda/projective_icp.hpp:8-46).
#pragma once
#include <tfusion/types.hpp>

namespace tfusion {
namespace cuda {

class KF_EXPORTS ProjectiveICP {
public:
    enum { MAX_PYRAMID_LEVELS = 4 };
    typedef std::vector<Depth> DepthPyr;
    typedef std::vector<Cloud> PointsPyr;
    typedef std::vector<Normals> NormalsPyr;

    ProjectiveICP();
    virtual ~ProjectiveICP();

    float getDistThreshold() const;
    void setDistThreshold(float distance);
    float getAngleThreshold() const;
    void setAngleThreshold(float angle);
    void setIterationsNum(const std::vector<int>& iters);
    int getUsedLevelsNum() const;
    const std::vector<int>& iterations() const { return iters_; }

    virtual bool estimateTransform(Affine3f& affine, const Intr& intr, const Frame& curr, const Frame& prev);
    virtual bool estimateTransform(Affine3f& affine, const Intr& intr, const DepthPyr& dcurr, const NormalsPyr ncurr,
                                   const DepthPyr dprev, const NormalsPyr nprev);
    // the variant the frame path uses: one cooperative launch for the whole coarse-to-fine loop, solve on the device
    virtual bool estimateTransform(Affine3f& affine, const Intr& intr, const PointsPyr& vcurr, const NormalsPyr ncurr,
                                   const PointsPyr vprev, const NormalsPyr nprev);

private:
    std::vector<int> iters_;
    float angle_thres_;
    float dist_thres_;
};

}  // namespace cuda
}  // namespace tfusion
