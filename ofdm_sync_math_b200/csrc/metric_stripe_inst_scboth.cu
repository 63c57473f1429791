// Explicit instantiations of the stripe kernel (metric_stripe.cuh): scboth.
#define OFS_STRIPE_INSTANTIATE
#include "metric_stripe.cuh"

namespace ofs {
OFS_STRIPE_FOR_KIND(OFS_STRIPE_DEFINE, OFS_SC_BOTH)
}  // namespace ofs
