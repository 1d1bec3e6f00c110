/* srp-b200 -- barycentric checkpoints for large triangles (sm_100a).
 *
 * The reference walks a triangle's bounding box with incrementally accumulated barycentrics
 * (src/raster/triangle.c:102-109): pixel (x, y) is reached by (y - minY) float additions of
 * dlambda/dy followed by (x - minX) additions of dlambda/dx.  A tile that lies far inside a
 * large triangle would have to replay thousands of those additions per pixel block.  For
 * triangles whose box exceeds SRPD_LARGE_EXTENT pixels this pre-pass walks every pixel row
 * ONCE -- one thread per row, exactly the reference's sequence of additions -- and stores the
 * value at the first covered pixel of every tile column the row crosses.  The tile kernel
 * then starts from the checkpoint of (its row, its tile column) and adds at most a tile's
 * width of x steps.  Bit-exact by construction: the same additions in the same order.
 *
 * Traffic: 12 bytes per (row, tile column) written here, read once per pixel block row. */
#include "kernels.cuh"

__global__ void __launch_bounds__(128)
srpdCheckpointKernel(const __grid_constant__ SrpdCkptArgs a)
{
	srpdGridDependencyEnter();
	uint32_t nLarge = *a.largeCount;
	if (nLarge > a.largeCapacity)
		nLarge = a.largeCapacity;
	for (uint32_t i = blockIdx.x; i < nLarge; i += gridDim.x)
	{
		const uint2 item = a.largeList[i];
		const uint4* h = (const uint4*) (a.records + ((size_t) item.x * a.recCapacity + item.y) * a.recStride);
		const uint4 q0 = __ldg(h + 0), q1 = __ldg(h + 1), q2 = __ldg(h + 2), q4 = __ldg(h + 4);
		if (q4.w == 0)
			continue;
		const int minX = (int) (q0.w & 0xFFFFu), maxX = (int) (q0.w >> 16);
		const int minY = (int) (q1.w & 0xFFFFu), maxY = (int) (q1.w >> 16);
		const int rows = maxY - minY;
		const int col0 = minX / SRPD_TILE_W;
		const int cols = (maxX - 1) / SRPD_TILE_W - col0 + 1;
		const float dx0 = __uint_as_float(q1.x), dx1 = __uint_as_float(q1.y), dx2 = __uint_as_float(q1.z);
		const float dy0 = __uint_as_float(q2.x), dy1 = __uint_as_float(q2.y), dy2 = __uint_as_float(q2.z);
		float* table = a.ckptTable + 3 * (size_t) (q4.w - 1);
		for (int r = threadIdx.x; r < rows; r += blockDim.x)
		{
			float l0 = __uint_as_float(q0.x), l1 = __uint_as_float(q0.y), l2 = __uint_as_float(q0.z);
			for (int k = 0; k < r; k++)
			{
				l0 = __fadd_rn(l0, dy0); l1 = __fadd_rn(l1, dy1); l2 = __fadd_rn(l2, dy2);
			}
			int x = minX;
			float* out = table + 3 * (size_t) r * cols;
			for (int c = 0; c < cols; c++)
			{
				out[3 * c + 0] = l0; out[3 * c + 1] = l1; out[3 * c + 2] = l2;
				const int next = (col0 + c + 1) * SRPD_TILE_W;
				const int stop = next < maxX ? next : maxX;
				for (; x < stop; x++)
				{
					l0 = __fadd_rn(l0, dx0); l1 = __fadd_rn(l1, dx1); l2 = __fadd_rn(l2, dx2);
				}
			}
		}
	}
}

void srpdLaunchCheckpoints(const SrpdCkptArgs& a, cudaStream_t stream)
{
	srpdLaunchKernel(srpdCheckpointKernel, 296, 128, 0, stream, a);
}
