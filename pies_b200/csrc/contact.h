// contact.h — collision response kernels (contact.cu).
#pragma once

#include <algorithm>

#include "engine.h"

namespace pies {

struct ContactWork {
  DevBuf<uint32_t> sweepCounters;  // one chunk counter per ordered sweep of the substep
  DevBuf<float4> contribC;         // 4 per point-triangle entry
  uint32_t nTri = 0, sweepsUsed = 0;
};

// Once per substep, after detection: resets the chunk counters of the ordered sweeps.
int prepareContactSweeps(ContactWork& w, cudaStream_t s, const ContactLists& c);
// Per PD iteration: collision projections (local step) and their RHS contributions.
int launchContactProject(cudaStream_t s, const ContactLists& c, const float4* q, float thickness, float4* contribC,
                         float4* snap);
int launchGatherContacts(cudaStream_t s, uint32_t n, const ContactLists& c, const float4* contribC, const float4* snap,
                         float4* rhs);
// End of substep: ordered stabilisation sweeps (list order per node, see contact.cu) and the friction pass.
int launchStabilize(cudaStream_t s, ContactWork& w, const ContactLists& c, uint32_t n, float4* q, float4* prev,
                    const float4* snap, float thickness, uint32_t iterations);
int launchFriction(cudaStream_t s, ContactWork& w, const ContactLists& c, uint32_t n, const float4* q, float4* vel,
                   float friction, float staticThreshold);

}  // namespace pies
