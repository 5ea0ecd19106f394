// Host-side micro-benchmark of the 2-bit read packer (cammiq_b200/csrc/pack_reads.cpp): thread
// scaling of packBatch against a pure streaming pass over the same bytes.  Build and run:
//   g++ -O3 -std=c++11 -pthread -Icammiq_b200/csrc tools/packbench.cpp cammiq_b200/csrc/pack_reads.cpp -o /tmp/packbench && /tmp/packbench
#include "pack_reads.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
using namespace cammiq;
int main() {
	uint64_t n = 10000000; uint32_t rl = 100;
	std::vector<uint8_t> bases(n * rl), lengths(n, rl), out(n * 25 + 64), ol(n);
	for (size_t i = 0; i < bases.size(); i++) bases[i] = "ACGT"[(i * 2654435761u >> 13) & 3];
	printf("isa %s\n", packIsaName());
	for (int T : {1, 2, 4, 8, 12, 16}) {
		WorkerPool pool(T);
		AsciiReads in = {bases.data(), NULL, rl, lengths.data()};
		double best = 1e9, bestcp = 1e9;
		for (int rep = 0; rep < 5; rep++) {
			auto t0 = std::chrono::high_resolution_clock::now();
			PackedLayout L = planBatch(pool, lengths.data(), 0, n, false);
			packBatch(pool, in, 0, n, L, out.data(), NULL, ol.data());
			best = std::min(best, std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count());
			// pure streaming of the same bytes: read 1 GB, write 0.25 GB
			t0 = std::chrono::high_resolution_clock::now();
			pool.run([&](int t) {
				uint64_t per = n / T, a = per * t, b = t == T - 1 ? n : a + per;
				uint64_t acc = 0; const uint64_t *p = (const uint64_t *) (bases.data() + a * rl); uint64_t words = (b - a) * rl / 8;
				uint64_t *q = (uint64_t *) (out.data() + (a * 25 & ~7ull));
				for (uint64_t i = 0; i + 4 <= words; i += 4) { acc += p[i] ^ p[i+1] ^ p[i+2] ^ p[i+3]; q[i >> 2] = acc; }
				ol[a] = (uint8_t) acc;
			});
			bestcp = std::min(bestcp, std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count());
		}
		printf("T=%2d pack %.1f ms (%.0f M reads/s, %.1f GB/s in)   stream-only %.1f ms (%.1f GB/s in)\n", T, best, n / best / 1e3, n * rl / best / 1e6, bestcp, n * rl / bestcp / 1e6);
	}
}
