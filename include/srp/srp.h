/* srp-b200 -- umbrella header, same role as the reference's include/srp/srp.h:
 * the only header a program needs.  Define SRP_INCLUDE_VEC / SRP_INCLUDE_MAT before
 * including it to also get the vec2/3/4 and mat4 helpers. */
#pragma once
#include "srp/api.h"
#ifdef SRP_INCLUDE_VEC
	#include "srp/vec.h"
#endif
#ifdef SRP_INCLUDE_MAT
	#include "srp/mat.h"
#endif
