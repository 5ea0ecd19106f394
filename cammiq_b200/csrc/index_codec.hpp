// Host-side codec for CAMMiQ index files (index_u.bin1 / index_d.bin2 and their .aux).
// Format: SURVEY.md section 5.9; reference reader binaryio.cpp:141-214, hashtrie.cpp:425-507,
// reference writer binaryio.cpp:3-134, hashtrie.cpp:595-700.  Written from the format, not
// from the reference code: iterative bit-stream decoder into flat arrays, no per-node heap
// objects.
#ifndef CAMMIQ_INDEX_CODEC_HPP
#define CAMMIQ_INDEX_CODEC_HPP

#include <cstdint>
#include <memory>
#include <new>
#include <string>
#include <vector>

namespace cammiq {

// Reference encoding shared by bucket roots and trie children.
//   0                      no child / no trie
//   0x80000000 | leaf_id   a leaf (file-order id)
//   node_id + 1            an internal node; nodes[4*node_id + code] holds its children
static const uint32_t kRefNone = 0u;
static const uint32_t kRefLeafTag = 0x80000000u;
inline bool refIsLeaf(uint32_t r) { return (r & kRefLeafTag) != 0; }
inline uint32_t refLeafId(uint32_t r) { return r & ~kRefLeafTag; }
inline uint32_t refNodeId(uint32_t r) { return r - 1; }

// std::vector whose resize() leaves new elements uninitialised: the decoder sizes its arrays
// once and lets the worker threads that fill them take the page faults.
template <class T>
struct NoInitAlloc : std::allocator<T> {
	template <class U> struct rebind { typedef NoInitAlloc<U> other; };
	NoInitAlloc() {}
	template <class U> NoInitAlloc(const NoInitAlloc<U> &) {}
	template <class U> void construct(U *p) { ::new ((void *) p) U; }
	template <class U, class A> void construct(U *p, const A &a) { ::new ((void *) p) U(a); }
};
template <class T>
struct FlatVec {
	typedef std::vector<T, NoInitAlloc<T> > type;
};

struct DecodedIndex {
	bool doubly_unique = false;
	uint32_t hash_len = 0;
	// buckets in file order
	FlatVec<uint64_t>::type bucket_key;
	FlatVec<uint32_t>::type bucket_root;
	// internal trie nodes, 4 child refs each
	FlatVec<uint32_t>::type nodes;
	// leaves in file order
	FlatVec<uint32_t>::type ref_id1, ref_id2;
	FlatVec<uint16_t>::type ucount1, ucount2;
	FlatVec<uint8_t>::type depth;
	uint32_t max_ref_id = 0;

	uint64_t numLeaves() const { return ref_id1.size(); }
	uint64_t numNodes() const { return nodes.size() / 4; }
};

// Returns 0 or a CQ_E* code (include/cammiq_gpu.h); err receives a message.
int decodeIndexFile(const std::string &path, DecodedIndex &out, std::string &err);

// Writer for the same format (tooling: synthetic indices, round-trip tests).
// Buckets are emitted in the order given.
int encodeIndexFile(const std::string &path, const DecodedIndex &idx, std::string &err);

// 2-bit code of a base, -1 for anything outside ACGTacgt (query.cpp:1860-1883).
int baseCode(uint8_t c);

} // namespace cammiq
#endif
