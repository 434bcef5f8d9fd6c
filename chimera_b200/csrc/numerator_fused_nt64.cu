// numerator_fused_nt64.cu -- third instantiation of the fused 1-D kernel of numerator_fused.cu: 64 threads per CTA, twelve
// co-resident CTAs per SM (A/B candidate for very short events; see the note at the top of numerator_fused.cu).
#define FU_NT 64
#define CHB_FU_VARIANT _nt64
#define CHB_FU_MINB 12
#include "numerator_fused.cu"
