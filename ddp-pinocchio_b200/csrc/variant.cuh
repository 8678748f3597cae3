// variant.cuh -- joint-type variant of the device code of a translation unit.
//
// kernels_nj.cu is compiled twice per model size: DDPB_PRISMATIC = 0 (every joint revolute: the reference's arms, the bench
// workloads) and DDPB_PRISMATIC = 1 (some joint prismatic).  The joint-type tests of the rigid-body layer (rbd.cuh) are
// compile-time dead in the first variant, so all-revolute models run exactly the straight-line code they ran before
// prismatic joints existed.  Measured on B200 when the tests were runtime branches: per-joint tests inside the unrolled
// joint loops cost C2 +3 % and the rollouts of the arms +6 %; one branch per task with both bodies in the same kernel
// gave C2 back but cost the C5 linearization +4.4 % and the rollouts +3 % through code size alone (these kernels run
// 10-20 thousand straight-line instructions once per CTA and already stall 9-12 % of their samples on instruction fetch).
// Everything that depends on the variant lives in an inline namespace named after it, so the two variants of a function or
// kernel are different entities (no ODR games); the PODs shared with the host (ModelDev, BatchDev, Launch) stay in ddpb.
#pragma once
#ifndef DDPB_PRISMATIC
#define DDPB_PRISMATIC 0
#endif
#if DDPB_PRISMATIC
#define DDPB_VARIANT_NS kin_rp
#else
#define DDPB_VARIANT_NS kin_r
#endif
#define DDPB_VARIANT_BEGIN namespace ddpb { inline namespace DDPB_VARIANT_NS {
#define DDPB_VARIANT_END } }
namespace ddpb {
constexpr bool HAS_PRISMATIC = DDPB_PRISMATIC != 0;
}
