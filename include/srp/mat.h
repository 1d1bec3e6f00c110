/* srp-b200 -- mat4 helpers (API of the reference's include/srp/mat.h:16-93).
 * Row-major 4x4 float matrix.  See srp/vec.h for how the same names resolve on the
 * host (library symbols) and under nvcc (inline __host__ __device__ twins). */
#pragma once
#include "srp/vec.h"

typedef struct mat4 { float data[4][4]; } mat4;

#if !defined(__cplusplus)
/* A*b and A*B; every element is the left-associative sum of its four products */
vec4 mat4MultiplyVec4(const mat4* restrict a, vec4 b);
mat4 mat4MultiplyMat4(const mat4* restrict a, const mat4* restrict b);

mat4 mat4ConstructIdentity(void);
mat4 mat4ConstructScale(float x, float y, float z);
mat4 mat4ConstructTranslate(float x, float y, float z);
/* rotation by x, y, z radians about the X, Y, Z axes (double-precision sin/cos) */
mat4 mat4ConstructRotate(float x, float y, float z);
/* T * (R * S) */
mat4 mat4ConstructTRS(float transX, float transY, float transZ,
                      float rotationX, float rotationY, float rotationZ,
                      float scaleX, float scaleY, float scaleZ);
/* TRS of the negated camera position and rotation */
mat4 mat4ConstructView(float cameraX, float cameraY, float cameraZ,
                       float rotationX, float rotationY, float rotationZ,
                       float scaleX, float scaleY, float scaleZ);
/* maps the box [min, max] to the NDC cube */
mat4 mat4ConstructOrthogonalProjection(float x_min, float x_max, float y_min, float y_max,
                                       float z_min, float z_max);
/* orthogonal(box) * perspective(z_near, z_far) */
mat4 mat4ConstructPerspectiveProjection(float x_min_near, float x_max_near,
                                        float y_min_near, float y_max_near,
                                        float z_near, float z_far);
#else
	#include "srp/detail/mat_inline.h"
#endif
