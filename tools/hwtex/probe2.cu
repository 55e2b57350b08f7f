// Second probe: individual trilinear weights.  Texture A: level 0 = one texel 255 at (3,3) of 8x8, level 1 = 0.
// Texture B: level 0 = 0, level 1 = one texel 255 at (1,1) of 4x4.  Sampled over grids of (a, b, gamma) so that the
// output is the weight the unit gave that single texel.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if(e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while(0)
__global__ void k_sample(cudaTextureObject_t tex, const float *uvl, float4 *out, int n) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n)
		out[i] = tex2DLod<float4>(tex, uvl[i * 3], uvl[i * 3 + 1], uvl[i * 3 + 2]);
}
static cudaTextureObject_t makeTex(const std::vector<std::vector<unsigned char>> &levels, int w, int h) {
	cudaMipmappedArray_t arr;
	cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
	CK(cudaMallocMipmappedArray(&arr, &desc, make_cudaExtent(w, h, 0), (unsigned)levels.size()));
	for(size_t l = 0; l < levels.size(); l++) {
		int lw = w >> l > 0 ? w >> l : 1, lh = h >> l > 0 ? h >> l : 1;
		cudaArray_t a;
		CK(cudaGetMipmappedArrayLevel(&a, arr, (unsigned)l));
		CK(cudaMemcpy2DToArray(a, 0, 0, levels[l].data(), lw * 4, lw * 4, lh, cudaMemcpyHostToDevice));
	}
	cudaResourceDesc res{};
	res.resType = cudaResourceTypeMipmappedArray;
	res.res.mipmap.mipmap = arr;
	cudaTextureDesc td{};
	td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
	td.filterMode = cudaFilterModeLinear, td.mipmapFilterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeNormalizedFloat, td.normalizedCoords = 1;
	td.maxAnisotropy = 1, td.maxMipmapLevelClamp = float(levels.size() - 1);
	cudaTextureObject_t obj;
	CK(cudaCreateTextureObject(&obj, &res, &td, nullptr));
	return obj;
}
int main(int argc, char **argv) {
	FILE *f = fopen(argc > 1 ? argv[1] : "hwtex_probe2.bin", "wb");
	std::vector<unsigned char> a0(8 * 8 * 4, 0), a1(4 * 4 * 4, 0), b0(8 * 8 * 4, 0), b1(4 * 4 * 4, 0);
	for(int c = 0; c < 4; c++)
		a0[(3 * 8 + 3) * 4 + c] = 255, b1[(1 * 4 + 1) * 4 + c] = 255;
	cudaTextureObject_t ta = makeTex({a0, a1}, 8, 8), tb = makeTex({b0, b1}, 8, 8);
	// positions: level-0 texel coordinates x in [2.5, 3.5) -> weight a of texel 3 rises from 0 to 1 (texel (3,3) is the
	// "11" corner of the footprint); same for y.  In level 1 the same u, v fall at x1 = x / 2: in [1.25, 1.75) - 0.5 = [0.75, 1.25):
	// footprint (0,1) below 1.0 and (1,2) above.
	std::vector<float> uvl;
	for(int ia = 0; ia <= 256; ia += 4)
		for(int ib = 0; ib <= 256; ib += 4)
			for(int ig = 0; ig < 256; ig += 5) {
				uvl.push_back((2.5f + ia / 256.0f) / 8.0f), uvl.push_back((2.5f + ib / 256.0f) / 8.0f), uvl.push_back(ig / 256.0f);
			}
	int n = (int)uvl.size() / 3;
	float *d_in;
	float4 *d_out;
	CK(cudaMalloc(&d_in, uvl.size() * 4));
	CK(cudaMalloc(&d_out, n * 16));
	CK(cudaMemcpy(d_in, uvl.data(), uvl.size() * 4, cudaMemcpyHostToDevice));
	std::vector<float> out(n * 4);
	fwrite(&n, 4, 1, f);
	fwrite(uvl.data(), 4, uvl.size(), f);
	for(cudaTextureObject_t t : {ta, tb}) {
		k_sample<<<(n + 255) / 256, 256>>>(t, d_in, d_out, n);
		CK(cudaDeviceSynchronize());
		CK(cudaMemcpy(out.data(), d_out, n * 16, cudaMemcpyDeviceToHost));
		fwrite(out.data(), 4, out.size(), f);
	}
	fclose(f);
	printf("%d samples\n", n);
	return 0;
}
