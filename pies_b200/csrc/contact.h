// contact.h — collision response kernels (contact.cu).
#pragma once

#include <algorithm>

#include "engine.h"

namespace pies {

struct ContactWork {
  DevBuf<uint32_t> parent, perm, tmpVals, heads, compStart, nComp, sortHist, scanScratch;
  DevBuf<uint64_t> keys, tmpKeys;
  DevBuf<float4> contribC;  // 4 per point-triangle entry
  uint32_t nTri = 0;
};

// Once per substep, after detection: connected components of the contact graph + entry order per component.
int buildContactComponents(ContactWork& w, cudaStream_t s, uint32_t nNodes, const ContactLists& c);
// Per PD iteration: collision projections (local step) and their RHS contributions.
int launchContactProject(cudaStream_t s, const ContactLists& c, const float4* q, float thickness, float4* contribC,
                         float4* snap);
int launchGatherContacts(cudaStream_t s, uint32_t n, const ContactLists& c, const float4* contribC, const float4* snap,
                         float4* rhs);
// End of substep: ordered stabilisation sweeps and the friction pass.
int launchStabilize(cudaStream_t s, const ContactWork& w, const ContactLists& c, float4* q, float4* prev, const float4* snap,
                    float thickness, uint32_t iterations);
int launchFriction(cudaStream_t s, const ContactWork& w, const ContactLists& c, uint32_t n, const float4* q, float4* vel,
                   float friction, float staticThreshold);

}  // namespace pies
