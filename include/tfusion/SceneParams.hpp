// Scene parameters (reference: include/tfusion/SceneParams.hpp:45-53 — same members, same constructor order).
#pragma once
namespace tfusion {
class SceneParams {
public:
    float voxelSize;                         // metres
    float viewFrustum_min, viewFrustum_max;  // metres
    float mu;                                // truncation band, metres
    int maxW;                                // running-average cap
    bool stopIntegratingAtMaxW;
    SceneParams() {}
    SceneParams(float mu_, int maxW_, float voxelSize_, float viewFrustum_min_, float viewFrustum_max_, bool stopIntegratingAtMaxW_)
        : voxelSize(voxelSize_), viewFrustum_min(viewFrustum_min_), viewFrustum_max(viewFrustum_max_), mu(mu_), maxW(maxW_),
          stopIntegratingAtMaxW(stopIntegratingAtMaxW_) {}
    explicit SceneParams(const SceneParams* o) { SetFrom(o); }
    void SetFrom(const SceneParams* o) { *this = *o; }
};
}  // namespace tfusion
