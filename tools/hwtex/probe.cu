// Probe of the texture unit's filtering arithmetic (experiment behind LUCID_HW_TEXTURE): samples small RGBA8
// mipmapped textures through tex2DLod at controlled coordinates and dumps the raw float results, so that a CPU model
// of the filter (weight quantisation, rounding) can be fitted and verified offline.
//   nvcc -arch=sm_100a -o tools/hwtex/probe tools/hwtex/probe.cu && tools/hwtex/probe gpurun_out/hwtex_probe.bin
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

struct Tex {
	cudaMipmappedArray_t arr;
	cudaTextureObject_t obj;
};
static Tex makeTex(const std::vector<std::vector<unsigned char>> &levels, int w, int h) {
	Tex t;
	cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
	CK(cudaMallocMipmappedArray(&t.arr, &desc, make_cudaExtent(w, h, 0), (unsigned)levels.size()));
	for(size_t l = 0; l < levels.size(); l++) {
		int lw = w >> l > 0 ? w >> l : 1, lh = h >> l > 0 ? h >> l : 1;
		cudaArray_t a;
		CK(cudaGetMipmappedArrayLevel(&a, t.arr, (unsigned)l));
		CK(cudaMemcpy2DToArray(a, 0, 0, levels[l].data(), lw * 4, lw * 4, lh, cudaMemcpyHostToDevice));
	}
	cudaResourceDesc res{};
	res.resType = cudaResourceTypeMipmappedArray;
	res.res.mipmap.mipmap = t.arr;
	cudaTextureDesc td{};
	td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
	td.filterMode = cudaFilterModeLinear, td.mipmapFilterMode = cudaFilterModeLinear;
	td.readMode = cudaReadModeNormalizedFloat, td.normalizedCoords = 1;
	td.maxAnisotropy = 1, td.maxMipmapLevelClamp = float(levels.size() - 1);
	CK(cudaCreateTextureObject(&t.obj, &res, &td, nullptr));
	return t;
}
static unsigned rng_state = 12345;
static unsigned rnd() {
	rng_state = rng_state * 1664525u + 1013904223u;
	return rng_state >> 8;
}

int main(int argc, char **argv) {
	FILE *f = fopen(argc > 1 ? argv[1] : "hwtex_probe.bin", "wb");
	auto run = [&](const Tex &t, const std::vector<float> &uvl, const char *tag) {
		int n = (int)uvl.size() / 3;
		float *d_in;
		float4 *d_out;
		CK(cudaMalloc(&d_in, uvl.size() * 4));
		CK(cudaMalloc(&d_out, n * 16));
		CK(cudaMemcpy(d_in, uvl.data(), uvl.size() * 4, cudaMemcpyHostToDevice));
		k_sample<<<(n + 255) / 256, 256>>>(t.obj, d_in, d_out, n);
		CK(cudaDeviceSynchronize());
		std::vector<float> out(n * 4);
		CK(cudaMemcpy(out.data(), d_out, n * 16, cudaMemcpyDeviceToHost));
		char name[16] = {0};
		strncpy(name, tag, 15);
		fwrite(name, 1, 16, f);
		fwrite(&n, 4, 1, f);
		fwrite(uvl.data(), 4, uvl.size(), f);
		fwrite(out.data(), 4, out.size(), f);
		cudaFree(d_in), cudaFree(d_out);
		printf("%s: %d samples\n", tag, n);
	};
	// test 1: weight staircase in x on an 8x8 level: column 3 = 255, others 0 (all channels), all rows alike
	{
		std::vector<unsigned char> l0(8 * 8 * 4, 0);
		for(int y = 0; y < 8; y++)
			for(int c = 0; c < 4; c++)
				l0[(y * 8 + 3) * 4 + c] = 255;
		Tex t = makeTex({l0}, 8, 8);
		std::vector<float> uvl;
		for(int i = 0; i <= 8192; i++) {
			uvl.push_back((2.5f + i / 8192.0f) / 8.0f), uvl.push_back(2.5f / 8.0f), uvl.push_back(0.0f);
		}
		run(t, uvl, "stair_x8");
	}
	// test 2: the same staircase on a 4096-wide level (coordinate precision), column 1000 = 255
	{
		std::vector<unsigned char> l0((size_t)4096 * 4 * 4, 0);
		for(int y = 0; y < 4; y++)
			for(int c = 0; c < 4; c++)
				l0[((size_t)y * 4096 + 1000) * 4 + c] = 255;
		Tex t = makeTex({l0}, 4096, 4);
		std::vector<float> uvl;
		for(int i = 0; i <= 8192; i++) {
			uvl.push_back((999.5f + i / 8192.0f) / 4096.0f), uvl.push_back(0.5f / 4.0f), uvl.push_back(0.0f);
		}
		run(t, uvl, "stair_x4096");
	}
	// test 3: value rounding: two columns a | b for every pair (a, b) on a coarse grid, weights across
	{
		std::vector<float> uvl;
		std::vector<unsigned char> l0(512 * 2 * 4, 0);
		// 256 pairs laid out along x: texel 2k = a_k, 2k+1 = b_k
		for(int k = 0; k < 256; k++) {
			unsigned a = rnd() & 255, b = rnd() & 255;
			for(int y = 0; y < 2; y++)
				for(int c = 0; c < 4; c++) {
					l0[((size_t)y * 512 + 2 * k) * 4 + c] = (unsigned char)(c == 0 ? a : c == 1 ? b : c == 2 ? (a ^ 0x55) : k);
					l0[((size_t)y * 512 + 2 * k + 1) * 4 + c] = (unsigned char)(c == 0 ? b : c == 1 ? a : c == 2 ? (b ^ 0xaa) : 255 - k);
				}
		}
		Tex t = makeTex({l0}, 512, 2);
		for(int k = 0; k < 256; k++)
			for(int i = 0; i <= 256; i++) {
				uvl.push_back((2 * k + 0.5f + i / 256.0f) / 512.0f), uvl.push_back(0.25f), uvl.push_back(0.0f);
			}
		fwrite("texdata3", 1, 8, f);
		fwrite(l0.data(), 1, l0.size(), f);
		run(t, uvl, "values_x");
	}
	// test 4: bilinear: random 64x64 texture, random positions
	std::vector<std::vector<unsigned char>> chain;
	{
		int w = 64;
		for(int l = 0; l < 4; l++, w >>= 1) {
			std::vector<unsigned char> lv((size_t)w * w * 4);
			for(auto &b : lv)
				b = (unsigned char)(rnd() & 255);
			chain.push_back(lv);
		}
		Tex t = makeTex(chain, 64, 64);
		std::vector<float> uvl;
		for(int i = 0; i < 200000; i++) {
			uvl.push_back((rnd() & 0xffffff) / 16777216.0f * 3.0f - 1.0f), uvl.push_back((rnd() & 0xffffff) / 16777216.0f * 3.0f - 1.0f),
				uvl.push_back(0.0f);
		}
		fwrite("texdata4", 1, 8, f);
		for(auto &lv : chain)
			fwrite(lv.data(), 1, lv.size(), f);
		run(t, uvl, "bilinear");
		// test 5: trilinear on the same chain
		std::vector<float> uvl2;
		for(int i = 0; i < 200000; i++) {
			uvl2.push_back((rnd() & 0xffffff) / 16777216.0f), uvl2.push_back((rnd() & 0xffffff) / 16777216.0f),
				uvl2.push_back((rnd() & 0xffffff) / 16777216.0f * 3.5f - 0.25f);
		}
		run(t, uvl2, "trilinear");
	}
	// test 6: lod staircase: level 0 all 0, level 1 all 255
	{
		std::vector<unsigned char> l0(8 * 8 * 4, 0), l1(4 * 4 * 4, 255);
		Tex t = makeTex({l0, l1}, 8, 8);
		std::vector<float> uvl;
		for(int i = 0; i <= 8192; i++) {
			uvl.push_back(0.3f), uvl.push_back(0.3f), uvl.push_back(i / 8192.0f);
		}
		run(t, uvl, "stair_lod");
	}
	fclose(f);
	return 0;
}
