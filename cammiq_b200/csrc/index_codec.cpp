#include "index_codec.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/cammiq_gpu.h"

namespace cammiq {

int baseCode(uint8_t c) {
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return -1;
	}
}

namespace {

// MSB-first bit cursor over the AUX stream; reads past the end return 1 (the reference
// reader yields all-ones there, binaryio.cpp:146-149).  A 64-bit window holds the next bits
// left-aligned and is refilled eight bytes at a time.
struct BitCursor {
	const uint8_t *p;
	uint64_t nbits, pos; // pos = stream position of the window's first bit
	uint64_t win = 0;
	uint32_t have = 0;   // valid bits in win
	BitCursor(const uint8_t *data, uint64_t n_bits, uint64_t at) : p(data), nbits(n_bits), pos(at) {}
	inline void refill() {
		const uint64_t byte = pos >> 3, nbytes = (nbits + 7) >> 3;
		uint64_t v;
		if (byte + 8 <= nbytes) {
			memcpy(&v, p + byte, 8);
			v = __builtin_bswap64(v);
		} else {
			v = 0;
			for (uint64_t i = 0; i < 8; i++)
				v = (v << 8) | (byte + i < nbytes ? p[byte + i] : 0xFFu);
		}
		const uint32_t skip = (uint32_t) (pos & 7);
		win = skip ? (v << skip) | ((1ull << skip) - 1) : v; // the vacated low bits are never consumed
		have = 64 - skip;
	}
	inline void skip(uint32_t n) {
		win <<= n;
		have -= n;
		pos += n;
	}
	inline uint32_t bit() {
		if (have < 1)
			refill();
		const uint32_t v = (uint32_t) (win >> 63);
		skip(1);
		return v;
	}
	inline uint32_t bits(int n) {
		uint32_t v = 0;
		for (int i = 0; i < n; i++)
			v = (v << 1) | bit();
		return v;
	}
	// next five bits without consuming; 0b10000 is a leaf (padding past the end is all ones, so
	// a truncated pattern can never read as one)
	inline uint32_t peek5() {
		if (have < 5)
			refill();
		return (uint32_t) (win >> 59);
	}
};

struct Frame {
	uint32_t node;  // id of this node (provisional until a child turns up)
	uint8_t next;   // next child slot to decode
	uint8_t depth;  // trie depth of this node (uint8 arithmetic like the reference)
	bool any;       // some child was non-NULL
};

// A whole file mapped read-only (falls back to reading it when it cannot be mapped).
struct FileView {
	const uint8_t *p = NULL;
	uint64_t n = 0;
	bool mapped = false;
	std::vector<uint8_t> owned;
	~FileView() {
		if (mapped && p != NULL)
			munmap((void *) p, n);
	}
	bool open(const std::string &fn) {
		int fd = ::open(fn.c_str(), O_RDONLY);
		if (fd < 0)
			return false;
		struct stat st;
		if (fstat(fd, &st) != 0) {
			close(fd);
			return false;
		}
		n = (uint64_t) st.st_size;
		if (n > 0) {
			void *m = mmap(NULL, n, PROT_READ, MAP_PRIVATE, fd, 0);
			if (m != MAP_FAILED) {
				p = (const uint8_t *) m;
				mapped = true;
				madvise(m, n, MADV_WILLNEED);
			} else {
				owned.resize(n);
				uint64_t got = 0;
				while (got < n) {
					ssize_t r = read(fd, owned.data() + got, n - got);
					if (r <= 0)
						break;
					got += (uint64_t) r;
				}
				if (got != n) {
					close(fd);
					return false;
				}
				p = owned.data();
			}
		}
		close(fd);
		return true;
	}
};

// The trie shape of one bucket is a pre-order bit string: '0' = no child, '1' + four children =
// a node, and a node whose four children are all absent is a leaf (hashtrie.cpp:432-457).  The
// walker below fills the flat arrays of a bucket range through StoreSink (pass 2, all threads);
// pass 1 only counts (scanBucketShape).
struct StoreSink {
	DecodedIndex *out;
	const uint8_t *ints; // INT stream
	uint64_t int_size;
	uint64_t bucket;     // bucket being decoded: its leaf records follow its key
	uint64_t leaves, nodes;
	uint32_t leaf_bytes, h;
	bool dd, bad_pair = false;
	uint32_t max_ref = 0;
	inline uint32_t newNode() {
		uint32_t *c = &out->nodes[4 * (size_t) nodes];
		c[0] = c[1] = c[2] = c[3] = kRefNone;
		return (uint32_t) nodes++;
	}
	inline void dropLastNode() { nodes--; }
	inline uint32_t leaf(uint8_t depth) {
		const uint64_t id = leaves++;
		const uint8_t *r = ints + 8 * (bucket + 1) + (uint64_t) leaf_bytes * id; // pass 1 checked the extent
		uint32_t r1, r2 = 0;
		uint16_t c1, c2 = 0;
		memcpy(&r1, r, 4);
		r1 = __builtin_bswap32(r1);
		if (dd) {
			memcpy(&r2, r + 4, 4);
			r2 = __builtin_bswap32(r2);
			c1 = (uint16_t) ((r[8] << 8) | r[9]);
			c2 = (uint16_t) ((r[10] << 8) | r[11]);
			bad_pair |= r1 == 0 || r2 == 0;
		} else
			c1 = (uint16_t) ((r[4] << 8) | r[5]);
		out->ref_id1[id] = r1;
		out->ref_id2[id] = r2;
		out->ucount1[id] = c1;
		out->ucount2[id] = c2;
		out->depth[id] = (uint8_t) (depth + h);
		max_ref = std::max(max_ref, std::max(r1, r2));
		return kRefLeafTag | (uint32_t) id;
	}
	inline void setChild(uint32_t node, uint8_t slot, uint32_t ref) { out->nodes[4 * (size_t) node + slot] = ref; }
};

// Decodes the shape of one bucket; returns its root reference, or false on a runaway stream.
template <class Sink>
inline bool walkBucket(BitCursor &bc, Sink &sk, std::vector<Frame> &stack, uint32_t &root) {
	root = kRefNone;
	// fast path: the bucket is one leaf at depth h ("10000")
	if (bc.peek5() == 0x10u) {
		bc.skip(5);
		root = sk.leaf(0);
		return true;
	}
	if (bc.bit() == 0)
		return true;
	stack.clear();
	Frame f0 = {sk.newNode(), 0, 0, false};
	stack.push_back(f0);
	while (!stack.empty()) {
		Frame &f = stack.back();
		if (f.next < 4) {
			uint8_t slot = f.next++;
			if (bc.peek5() == 0x10u) {
				bc.skip(5);
				f.any = true;
				sk.setChild(f.node, slot, sk.leaf((uint8_t) (f.depth + 1)));
			} else if (bc.bit() != 0) {
				if (stack.size() > 4096 || bc.pos > bc.nbits + 64)
					return false;
				f.any = true;
				const uint8_t d1 = (uint8_t) (f.depth + 1);
				const uint32_t parent = f.node, id = sk.newNode();
				sk.setChild(parent, slot, id + 1);
				Frame nf = {id, 0, d1, false};
				stack.push_back(nf); // invalidates f
			}
			continue;
		}
		// all four children decoded
		Frame done = f;
		stack.pop_back();
		if (!done.any) {
			// a node with four NULL children is a leaf; it is the most recently created node, so
			// it can be dropped from the node array
			sk.dropLastNode();
			uint32_t lr = sk.leaf(done.depth);
			if (stack.empty())
				root = lr;
			else {
				Frame &p = stack.back();
				sk.setChild(p.node, (uint8_t) (p.next - 1), lr);
			}
		} else if (stack.empty())
			root = done.node + 1;
	}
	return true;
}

// Pass 1 needs counts only, and a leaf is recognisable where it starts ('1' followed by four
// absent children), so the shape of a bucket can be scanned without a stack: `pending` child
// slots remain to be read; an absent child uses one up, a leaf uses one up, an internal node
// uses one up and opens four.  Internal nodes are counted in pre-order, which is the order the
// fill pass numbers them in.
inline bool scanBucketShape(BitCursor &bc, uint64_t &leaves, uint64_t &nodes) {
	uint32_t v = bc.peek5();
	if (v == 0x10u) {
		bc.skip(5);
		leaves++;
		return true;
	}
	bc.skip(1);
	if (!(v & 0x10u))
		return true; // no trie under this bucket
	nodes++;
	uint64_t pending = 4;
	while (pending > 0) {
		v = bc.peek5();
		if (!(v & 0x10u)) {
			// a run of absent children (up to the five bits in view)
			const uint64_t zeros = (uint64_t) __builtin_clz((v << 27) | (1u << 26));
			const uint64_t n = zeros < pending ? zeros : pending;
			bc.skip((uint32_t) n);
			pending -= n;
		} else if (v == 0x10u) {
			bc.skip(5);
			leaves++;
			pending--;
		} else {
			bc.skip(1);
			nodes++;
			pending += 3;
			if (pending > 3 * 4096 || bc.pos > bc.nbits + 64)
				return false;
		}
	}
	return true;
}

struct Checkpoint {
	uint64_t aux_pos, leaves, nodes;
};
static const uint64_t kCheckpointBuckets = 4096;

unsigned decodeThreads() {
	const char *env = getenv("CAMMIQ_DECODE_THREADS");
	unsigned n = env ? (unsigned) atoi(env) : std::thread::hardware_concurrency();
	return std::max(1u, std::min(n, 16u));
}

} // namespace

// Two passes.  Pass 1 walks the AUX bit stream alone on one thread (the stream is strictly
// sequential) and records, every 4096 buckets, the bit position and the number of leaves and
// internal nodes so far; that also fixes every bucket's position in the INT stream
// (8 bytes of key per bucket + one fixed-size record per leaf).  Pass 2 decodes the bucket
// ranges between checkpoints on all threads straight into the pre-sized flat arrays; ids are
// file order either way.
int decodeIndexFile(const std::string &path, DecodedIndex &out, std::string &err) {
	FileView ints, aux;
	if (!ints.open(path)) {
		err = "Cannot open file: " + path + ".";
		return CQ_EIO;
	}
	if (!aux.open(path + ".aux")) {
		err = "Cannot open file: " + path + ".aux.";
		return CQ_EIO;
	}
	BitCursor bc(aux.p, aux.n * 8, 0);
	out = DecodedIndex();
	out.doubly_unique = bc.bit() != 0;
	uint32_t option = bc.bits(7);
	out.hash_len = bc.bits(8);
	if (option != 64) {
		err = "Index " + path + ": header option is not 64.";
		return CQ_EFORMAT;
	}
	if (out.hash_len < 1 || out.hash_len > 31) {
		err = "Index " + path + ": hash length outside [1, 31].";
		return CQ_EFORMAT;
	}
	const bool dd = out.doubly_unique;
	const uint32_t h = out.hash_len, leaf_bytes = dd ? 12 : 6;

	auto t_start = std::chrono::high_resolution_clock::now();
	// the shape pass is one thread walking both mapped files front to back; without this it
	// would also take every page fault of the mappings by itself
	{
		const unsigned Tp = decodeThreads();
		std::vector<std::thread> pool;
		std::atomic<uint64_t> sink(0);
		for (unsigned t = 0; t < Tp; t++)
			pool.emplace_back([&, t]() {
				uint64_t acc = 0;
				const FileView *files[2] = {&ints, &aux};
				for (int f = 0; f < 2; f++) {
					const uint64_t n = files[f]->n, lo = n / Tp * t, hi = t + 1 == Tp ? n : n / Tp * (t + 1);
					for (uint64_t i = lo; i < hi; i += 4096)
						acc += files[f]->p[i];
				}
				sink += acc;
			});
		for (auto &th : pool) th.join();
	}
	auto t_touch = std::chrono::high_resolution_clock::now();
	// ---- pass 1: shape only
	std::vector<Checkpoint> cps;
	cps.reserve((size_t) (ints.n / (8 + leaf_bytes) / kCheckpointBuckets + 2));
	struct { uint64_t leaves, nodes; } cs = {0, 0};
	uint64_t buckets = 0;
	for (;;) {
		const uint64_t at = 8 * buckets + (uint64_t) leaf_bytes * cs.leaves;
		if (at + 8 > ints.n) {
			err = "Index " + path + ": INT stream ends before the END64 terminator.";
			return CQ_EFORMAT;
		}
		uint64_t key;
		memcpy(&key, ints.p + at, 8);
		if (key == UINT64_MAX)
			break;
		if (buckets % kCheckpointBuckets == 0) {
			Checkpoint c = {bc.pos, cs.leaves, cs.nodes};
			cps.push_back(c);
		}
		if (!scanBucketShape(bc, cs.leaves, cs.nodes) || bc.pos > bc.nbits + 64) {
			err = "Index " + path + ": AUX stream is malformed (runaway trie).";
			return CQ_EFORMAT;
		}
		buckets++;
		if (cs.leaves >= 0x7FFFFFF0ull || cs.nodes >= 0x7FFFFFF0ull) {
			err = "Index " + path + ": more than 2^31 leaves or nodes.";
			return CQ_EFORMAT;
		}
	}
	const uint64_t n_leaves = cs.leaves, n_nodes = cs.nodes;
	auto t_pass1 = std::chrono::high_resolution_clock::now();
	out.bucket_key.resize((size_t) buckets);
	out.bucket_root.resize((size_t) buckets);
	out.nodes.resize((size_t) n_nodes * 4);
	out.ref_id1.resize((size_t) n_leaves);
	out.ref_id2.resize((size_t) n_leaves);
	out.ucount1.resize((size_t) n_leaves);
	out.ucount2.resize((size_t) n_leaves);
	out.depth.resize((size_t) n_leaves);

	// ---- pass 2: fill, one checkpoint range at a time per thread
	const unsigned T = (unsigned) std::min<uint64_t>(decodeThreads(), std::max<uint64_t>(1, cps.size()));
	std::atomic<size_t> next(0);
	std::vector<uint32_t> max_ref(T, 0);
	std::atomic<bool> bad_pair(false), broken(false);
	auto work = [&](unsigned t) {
		std::vector<Frame> st;
		for (;;) {
			const size_t c = next.fetch_add(1);
			if (c >= cps.size())
				break;
			BitCursor b2(aux.p, aux.n * 8, cps[c].aux_pos);
			StoreSink sk;
			sk.out = &out;
			sk.ints = ints.p;
			sk.int_size = ints.n;
			sk.leaves = cps[c].leaves;
			sk.nodes = cps[c].nodes;
			sk.leaf_bytes = leaf_bytes;
			sk.h = h;
			sk.dd = dd;
			const uint64_t lo = (uint64_t) c * kCheckpointBuckets, hi = std::min(buckets, lo + kCheckpointBuckets);
			for (uint64_t b = lo; b < hi; b++) {
				uint64_t key;
				memcpy(&key, ints.p + 8 * b + (uint64_t) leaf_bytes * sk.leaves, 8);
				sk.bucket = b;
				uint32_t root;
				if (!walkBucket(b2, sk, st, root)) {
					broken = true;
					return;
				}
				out.bucket_key[(size_t) b] = __builtin_bswap64(key);
				out.bucket_root[(size_t) b] = root;
			}
			max_ref[t] = std::max(max_ref[t], sk.max_ref);
			if (sk.bad_pair)
				bad_pair = true;
		}
	};
	{
		std::vector<std::thread> pool;
		for (unsigned t = 1; t < T; t++)
			pool.emplace_back(work, t);
		work(0);
		for (auto &th : pool) th.join();
	}
	if (getenv("CAMMIQ_VERBOSE"))
		fprintf(stderr, "[decode] %s: %lu buckets, page-in %.0f ms, shape pass %.0f ms, fill pass %.0f ms on %u threads\n", path.c_str(),
			(unsigned long) buckets, std::chrono::duration<double, std::milli>(t_touch - t_start).count(),
			std::chrono::duration<double, std::milli>(t_pass1 - t_touch).count(),
			std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t_pass1).count(), T);
	if (broken) {
		err = "Index " + path + ": AUX stream is malformed (runaway trie).";
		return CQ_EFORMAT;
	}
	for (unsigned t = 0; t < T; t++)
		out.max_ref_id = std::max(out.max_ref_id, max_ref[t]);
	if (dd && bad_pair) {
		err = "Index " + path + ": doubly-unique leaf without two genome ids.";
		return CQ_EFORMAT;
	}
	return CQ_OK;
}

namespace {

struct BitSink {
	std::vector<uint8_t> bytes;
	uint32_t cur = 0;
	int nbits = 0;
	inline void bit(uint32_t b) {
		cur = (cur << 1) | (b & 1u);
		if (++nbits == 8) {
			bytes.push_back((uint8_t) cur);
			cur = 0;
			nbits = 0;
		}
	}
	inline void bits(int n, uint32_t v) {
		for (int i = n - 1; i >= 0; i--)
			bit((v >> i) & 1u);
	}
};

inline void putBE(std::vector<uint8_t> &o, uint64_t v, int nbytes) {
	for (int i = nbytes - 1; i >= 0; i--)
		o.push_back((uint8_t) (v >> (8 * i)));
}

} // namespace

int encodeIndexFile(const std::string &path, const DecodedIndex &idx, std::string &err) {
	BitSink aux;
	std::vector<uint8_t> ints;
	ints.reserve(idx.numLeaves() * (idx.doubly_unique ? 20 : 14) + 16);
	aux.bit(idx.doubly_unique ? 1 : 0);
	aux.bits(7, 64);
	aux.bits(8, idx.hash_len);

	auto emitLeaf = [&](uint32_t leaf) {
		aux.bits(5, 0x10);
		putBE(ints, idx.ref_id1[leaf], 4);
		if (idx.doubly_unique) {
			putBE(ints, idx.ref_id2[leaf], 4);
			putBE(ints, idx.ucount1[leaf], 2);
			putBE(ints, idx.ucount2[leaf], 2);
		} else
			putBE(ints, idx.ucount1[leaf], 2);
	};

	std::vector<std::pair<uint32_t, int>> stack; // (node id, next child)
	for (size_t b = 0; b < idx.bucket_key.size(); b++) {
		putBE(ints, idx.bucket_key[b], 8);
		uint32_t root = idx.bucket_root[b];
		if (root == kRefNone) {
			aux.bit(0);
			continue;
		}
		if (refIsLeaf(root)) {
			emitLeaf(refLeafId(root));
			continue;
		}
		aux.bit(1);
		stack.clear();
		stack.push_back(std::make_pair(refNodeId(root), 0));
		while (!stack.empty()) {
			std::pair<uint32_t, int> &f = stack.back();
			if (f.second == 4) {
				stack.pop_back();
				continue;
			}
			uint32_t c = idx.nodes[4 * (size_t) f.first + f.second++];
			if (c == kRefNone)
				aux.bit(0);
			else if (refIsLeaf(c))
				emitLeaf(refLeafId(c));
			else {
				aux.bit(1);
				stack.push_back(std::make_pair(refNodeId(c), 0));
			}
		}
	}
	// trailers: 72 one-bits, END64 + 0xFFFF (binaryio.cpp:115-123)
	for (int i = 0; i < 72; i++)
		aux.bit(1);
	// a trailing partial byte is NOT flushed by the reference writer; the reader returns
	// ones past EOF, so dropping it is equivalent
	putBE(ints, UINT64_MAX, 8);
	putBE(ints, 0xFFFF, 2);

	FILE *f = fopen(path.c_str(), "wb");
	FILE *g = fopen((path + ".aux").c_str(), "wb");
	bool ok = f != NULL && g != NULL;
	if (ok)
		ok = fwrite(ints.data(), 1, ints.size(), f) == ints.size() &&
			fwrite(aux.bytes.data(), 1, aux.bytes.size(), g) == aux.bytes.size();
	if (f) fclose(f);
	if (g) fclose(g);
	if (!ok) {
		err = "Cannot write index file: " + path + ".";
		return CQ_EIO;
	}
	return CQ_OK;
}

} // namespace cammiq
