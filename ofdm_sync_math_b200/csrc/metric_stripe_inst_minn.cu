// Explicit instantiations of the stripe kernel (metric_stripe.cuh): minn.
#define OFS_STRIPE_INSTANTIATE
#include "metric_stripe.cuh"

namespace ofs {
OFS_STRIPE_FOR_KIND(OFS_STRIPE_DEFINE, OFS_MINN)
}  // namespace ofs
