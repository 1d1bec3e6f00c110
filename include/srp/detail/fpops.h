/* srp-b200 -- the four rounding-exact float operations everything parity-relevant is
 * written in.  On the device they are the round-to-nearest intrinsics, which the
 * compiler never contracts into FMAs whatever -fmad / LTO options are in effect; on
 * the host they are the plain operators (host objects are built without contraction,
 * like the reference: ISO C => -ffp-contract=off). */
#pragma once
#if defined(__CUDA_ARCH__)
	#define SRP_FMUL(a, b) __fmul_rn((a), (b))
	#define SRP_FADD(a, b) __fadd_rn((a), (b))
	#define SRP_FSUB(a, b) __fsub_rn((a), (b))
	#define SRP_FDIV(a, b) __fdiv_rn((a), (b))
	#define SRP_FSQRT(a)   __fsqrt_rn((a))
	#define SRP_DMUL(a, b) __dmul_rn((a), (b))
	#define SRP_DADD(a, b) __dadd_rn((a), (b))
	#define SRP_DSUB(a, b) __dsub_rn((a), (b))
	#define SRP_DDIV(a, b) __ddiv_rn((a), (b))
#else
	#include <math.h>
	#define SRP_FMUL(a, b) ((float) ((float) (a) * (float) (b)))
	#define SRP_FADD(a, b) ((float) ((float) (a) + (float) (b)))
	#define SRP_FSUB(a, b) ((float) ((float) (a) - (float) (b)))
	#define SRP_FDIV(a, b) ((float) ((float) (a) / (float) (b)))
	#define SRP_FSQRT(a)   (sqrtf((a)))
	#define SRP_DMUL(a, b) ((double) ((double) (a) * (double) (b)))
	#define SRP_DADD(a, b) ((double) ((double) (a) + (double) (b)))
	#define SRP_DSUB(a, b) ((double) ((double) (a) - (double) (b)))
	#define SRP_DDIV(a, b) ((double) ((double) (a) / (double) (b)))
#endif
#if defined(__CUDACC__)
	#define SRP_HD __host__ __device__ static __forceinline__
#else
	#define SRP_HD static inline
#endif
