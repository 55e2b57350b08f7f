// sync.cu -- frame hand-over of the bin-row split over NVLink without a collective: a device that finished
// its strip stores the frame number into a flag in the gathering device's memory (system-scope release), the
// gathering device waits for all flags with one warp, and a device may only store into the shared image again
// once the gathering device has released it.  All of it is stream-ordered device work: no host round trip and
// no NCCL kernel sits between two frames.
#include "common.cuh"

namespace lucid {

__device__ __forceinline__ u32 loadAcquireSys(const u32 *ptr) {
	u32 v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long globalTimerNs() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// after everything enqueued before it on the stream: *flag = value, ordered after those writes system-wide
__global__ void k_signal(u32 *flag, u32 value) {
	pdlEntry(); // the preceding grid (the frame's last kernel) has completed and its stores, peer ones included, are performed
	__threadfence_system();
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// returns when flags[i] has reached `value` (wrap-around compare) for every i < count; gives up after timeout_ns
// and raises bit 2 of *status (a peer that never signals must not hang the device)
__global__ void k_wait_flags(const u32 *flags, int count, u32 value, u32 *status, unsigned long long timeout_ns) {
	pdlEntry();
	const unsigned long long t0 = globalTimerNs();
	for(int i = threadIdx.x; i < count; i += blockDim.x) {
		while((int)(loadAcquireSys(flags + i) - value) < 0) {
			if(globalTimerNs() - t0 > timeout_ns) {
				if(status)
					*status = 4u;
				break;
			}
			__nanosleep(200);
		}
	}
	__threadfence_system();
}

void launchSignal(u32 *flag, u32 value, cudaStream_t stream) { launchPDL(k_signal, 1, 1, 0, stream, flag, value); }
void launchWaitFlags(const u32 *flags, int count, u32 value, u32 *status, unsigned long long timeout_ns, cudaStream_t stream) {
	launchPDL(k_wait_flags, 1, 32, 0, stream, flags, count, value, status, timeout_ns);
}

} // namespace lucid
