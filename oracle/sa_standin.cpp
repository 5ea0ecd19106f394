/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Suffix-array stand-in for parallel-divsufsort so
 * that the UNMODIFIED reference builder (gsa.cpp / build.cpp) can be linked here and
 * used to produce real index files for golden fixtures.  Induced-sorting construction
 * (SA-IS, Nong/Zhang/Chan 2009), written from the published algorithm.
 * Contract = libdivsufsort's: SA[0..n) are the start positions of the suffixes of T in
 * ascending lexicographic order, where a suffix that is a proper prefix of another
 * sorts first (implicit end sentinel smaller than every symbol).
 */
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include "divsufsort.h"

namespace {

template <typename S>
void induced_sort(const S *s, int64_t *SA, int64_t n, int64_t K) {
	/* s[n-1] is a unique, smallest symbol (0). */
	std::vector<bool> stype(n);
	stype[n - 1] = true;
	if (n >= 2) stype[n - 2] = false;
	for (int64_t i = n - 3; i >= 0; i--)
		stype[i] = (s[i] < s[i + 1]) || (s[i] == s[i + 1] && stype[i + 1]);
	auto is_lms = [&](int64_t i) { return i > 0 && stype[i] && !stype[i - 1]; };

	std::vector<int64_t> bkt(K);
	auto buckets = [&](bool ends) {
		std::fill(bkt.begin(), bkt.end(), 0);
		for (int64_t i = 0; i < n; i++) bkt[s[i]]++;
		int64_t sum = 0;
		for (int64_t c = 0; c < K; c++) {
			sum += bkt[c];
			bkt[c] = ends ? sum : sum - bkt[c];
		}
	};
	auto induce = [&]() {
		buckets(false);
		for (int64_t i = 0; i < n; i++) {
			int64_t j = SA[i] - 1;
			if (SA[i] > 0 && !stype[j]) SA[bkt[s[j]]++] = j;
		}
		buckets(true);
		for (int64_t i = n - 1; i >= 0; i--) {
			int64_t j = SA[i] - 1;
			if (SA[i] > 0 && stype[j]) SA[--bkt[s[j]]] = j;
		}
	};

	/* Pass 1: sort the LMS substrings. */
	buckets(true);
	for (int64_t i = 0; i < n; i++) SA[i] = -1;
	for (int64_t i = 1; i < n; i++)
		if (is_lms(i)) SA[--bkt[s[i]]] = i;
	induce();

	int64_t n1 = 0;
	for (int64_t i = 0; i < n; i++)
		if (is_lms(SA[i])) SA[n1++] = SA[i];
	for (int64_t i = n1; i < n; i++) SA[i] = -1;

	int64_t names = 0, prev = -1;
	for (int64_t i = 0; i < n1; i++) {
		int64_t pos = SA[i];
		bool differs = false;
		for (int64_t d = 0; d < n; d++) {
			if (prev == -1 || s[pos + d] != s[prev + d] || stype[pos + d] != stype[prev + d]) {
				differs = true;
				break;
			} else if (d > 0 && (is_lms(pos + d) || is_lms(prev + d)))
				break;
		}
		if (differs) {
			names++;
			prev = pos;
		}
		SA[n1 + pos / 2] = names - 1;
	}
	for (int64_t i = n - 1, j = n - 1; i >= n1; i--)
		if (SA[i] >= 0) SA[j--] = SA[i];

	/* Pass 2: order the LMS suffixes (recurse when names collide). */
	int64_t *SA1 = SA, *s1 = SA + n - n1;
	if (names < n1)
		induced_sort<int64_t>(s1, SA1, n1, names);
	else
		for (int64_t i = 0; i < n1; i++) SA1[s1[i]] = i;

	/* Pass 3: induce the full order from the sorted LMS suffixes. */
	buckets(true);
	for (int64_t i = 1, j = 0; i < n; i++)
		if (is_lms(i)) s1[j++] = i;
	for (int64_t i = 0; i < n1; i++) SA1[i] = s1[SA1[i]];
	for (int64_t i = n1; i < n; i++) SA[i] = -1;
	for (int64_t i = n1 - 1; i >= 0; i--) {
		int64_t j = SA[i];
		SA[i] = -1;
		SA[--bkt[s[j]]] = j;
	}
	induce();
}

} // namespace

int divsufsort(const uint8_t *T, int64_t *SA, int64_t n) {
	if (n < 0 || (n > 0 && (T == NULL || SA == NULL))) return -1;
	if (n == 0) return 0;
	if (n == 1) { SA[0] = 0; return 0; }
	/* Shift symbols by one and append the explicit sentinel 0. */
	std::vector<uint16_t> s(n + 1);
	for (int64_t i = 0; i < n; i++) s[i] = (uint16_t) T[i] + 1;
	s[n] = 0;
	std::vector<int64_t> work(n + 1);
	induced_sort<uint16_t>(s.data(), work.data(), n + 1, 257);
	/* work[0] is the sentinel suffix. */
	memcpy(SA, work.data() + 1, sizeof(int64_t) * n);
	return 0;
}

int sufcheck(const uint8_t *T, const int64_t *SA, int64_t n, bool verbose) {
	for (int64_t i = 1; i < n; i++) {
		int64_t a = SA[i - 1], b = SA[i];
		int64_t la = n - a, lb = n - b, l = la < lb ? la : lb;
		int c = memcmp(T + a, T + b, l);
		if (c > 0 || (c == 0 && la > lb)) {
			if (verbose) fprintf(stderr, "sufcheck: order violated at rank %ld.\n", (long) i);
			return -1;
		}
	}
	return 0;
}
