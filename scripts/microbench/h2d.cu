// Host -> device transfer rates that bound the end-to-end numbers of the big batches (config 3: 2 GB of sequences):
// pinned memcpyAsync in one piece / in 16 pieces from 16 threads, pageable -> pinned packing with 1..16 threads, and both
// together (the engine's pack + upload).  nvcc -O2 -o h2d h2d.cu -lpthread ; ./h2d [MB]
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 2048) << 20;
    char *pageable = (char*)malloc(bytes), *pinned, *dev;
    memset(pageable, 1, bytes);
    cudaHostAlloc(&pinned, bytes, cudaHostAllocDefault);
    memset(pinned, 2, bytes);
    cudaMalloc(&dev, bytes);
    cudaStream_t st;
    cudaStreamCreate(&st);
    for (int rep = 0; rep < 2; ++rep) {
        double t0 = now();
        cudaMemcpyAsync(dev, pinned, bytes, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        double t1 = now();
        printf("pinned -> device, one piece:            %6.1f ms  %5.1f GB/s\n", 1e3 * (t1 - t0), bytes / (t1 - t0) / 1e9);
    }
    for (int T : {1, 4, 8, 16}) {
        double t0 = now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() { memcpy(pinned + bytes * t / T, pageable + bytes * t / T, bytes * (t + 1) / T - bytes * t / T); });
        for (auto& x : th) x.join();
        double t1 = now();
        printf("pageable -> pinned, %2d threads:          %6.1f ms  %5.1f GB/s\n", T, 1e3 * (t1 - t0), bytes / (t1 - t0) / 1e9);
    }
    for (int T : {4, 16}) {
        double t0 = now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t]() {
                const size_t a = bytes * t / T, b = bytes * (t + 1) / T;
                const size_t piece = 8u << 20;
                for (size_t o = a; o < b; o += piece) {
                    const size_t n = std::min(piece, b - o);
                    memcpy(pinned + o, pageable + o, n);
                    cudaMemcpyAsync(dev + o, pinned + o, n, cudaMemcpyHostToDevice, st);
                }
            });
        for (auto& x : th) x.join();
        double t1 = now();
        cudaStreamSynchronize(st);
        double t2 = now();
        printf("pack + upload, %2d threads, 8 MB pieces:  %6.1f ms packed, %6.1f ms on the device  %5.1f GB/s\n", T, 1e3 * (t1 - t0),
               1e3 * (t2 - t0), bytes / (t2 - t0) / 1e9);
    }
    return 0;
}
