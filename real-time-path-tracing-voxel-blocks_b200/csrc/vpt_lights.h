// Local emissive lights: the host-built list (host/vpt_lights.cpp) and what the trace kernels get of it (vpt_wave.cu).
#pragma once
#include "../../include/vpt.h"
#include <stdint.h>
#include <vector>

namespace vpt {

struct LightList
{
    std::vector<VptLightInfo> lights;  // two triangles per exposed face of an emissive voxel
    std::vector<uint32_t> faceKeys;    // ascending (linear voxel << 3) | face; light index = 2 * position + triangle
    std::vector<VptAliasBin> alias;    // over luminance(radiance) * area
};
void buildLightList(const uint8_t *idsChunk, int cx, int cy, int cz, const VptMaterial *materials, int nMaterials, const uint16_t *blockToMaterial,
                    LightList &out);
// previous light id -> current light id (-1: the light is gone), Restir.h:60-75
void buildLightRemap(const std::vector<uint32_t> &prevKeys, const std::vector<uint32_t> &curKeys, std::vector<int> &prevToCur); // keys = LightList::faceKeys
void faceFrame(int face, int x, int y, int z, float *A, float *u, float *v);

// device view
struct LightView
{
    const VptLightInfo *lights;
    const VptAliasBin *alias;
    const uint32_t *faceKeys;
    const int *prevToCur;   // valid when stateDirty
    int numLights, numFaces, prevNumLights, stateDirty;
};

} // namespace vpt
