// numerator_fused_nt128.cu -- second instantiation of the fused 1-D kernel of numerator_fused.cu: 128 threads per CTA, six
// co-resident CTAs per SM, for events with few posterior samples (see the note at the top of numerator_fused.cu).
#define FU_NT 128
#define CHB_FU_VARIANT _nt128
#define CHB_FU_MINB 6
#include "numerator_fused.cu"
