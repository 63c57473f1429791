// Explicit instantiations of the stripe kernel (metric_stripe.cuh): mb_sc.
#define OFS_STRIPE_INSTANTIATE
#include "metric_stripe.cuh"

namespace ofs {
OFS_STRIPE_FOR_KIND_MB(OFS_STRIPE_DEFINE, OFS_SC)
OFS_STRIPE_FOR_KIND_MB(OFS_STRIPE_DEFINE, OFS_SC_BOTH)
}  // namespace ofs
