#include "index_codec.hpp"

#include <cstdio>
#include <cstring>

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

bool slurp(const std::string &fn, std::vector<uint8_t> &buf) {
	FILE *f = fopen(fn.c_str(), "rb");
	if (f == NULL)
		return false;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	buf.resize((size_t) n);
	bool ok = (n == 0) || fread(buf.data(), 1, (size_t) n, f) == (size_t) n;
	fclose(f);
	return ok;
}

// MSB-first bit cursor over the AUX stream; reads past the end return 1 (the reference
// reader yields all-ones there, binaryio.cpp:146-149).
struct BitCursor {
	const uint8_t *p;
	uint64_t nbits, pos;
	inline uint32_t bit() {
		if (pos >= nbits) {
			pos++;
			return 1;
		}
		uint32_t v = (p[pos >> 3] >> (7 - (pos & 7))) & 1u;
		pos++;
		return v;
	}
	inline uint32_t bits(int n) {
		uint32_t v = 0;
		for (int i = 0; i < n; i++)
			v = (v << 1) | bit();
		return v;
	}
	// next five bits without consuming; 0b10000 is a leaf
	inline uint32_t peek5() {
		if (pos + 5 > nbits)
			return 0xFFu;
		uint64_t byte = pos >> 3;
		uint32_t w = ((uint32_t) p[byte] << 8) | (byte + 1 < ((nbits + 7) >> 3) ? p[byte + 1] : 0xFFu);
		return (w >> (11 - (pos & 7))) & 0x1Fu;
	}
};

struct ByteCursor {
	const uint8_t *p;
	uint64_t n, pos;
	bool overrun;
	// fast paths for the common widths (unaligned load + byte swap), with the same overrun rule
	inline uint64_t be8() {
		if (pos + 8 <= n) {
			uint64_t v;
			memcpy(&v, p + pos, 8);
			pos += 8;
			return __builtin_bswap64(v);
		}
		return be(8);
	}
	inline uint32_t be4() {
		if (pos + 4 <= n) {
			uint32_t v;
			memcpy(&v, p + pos, 4);
			pos += 4;
			return __builtin_bswap32(v);
		}
		return (uint32_t) be(4);
	}
	inline uint16_t be2() {
		if (pos + 2 <= n) {
			uint16_t v = (uint16_t) ((p[pos] << 8) | p[pos + 1]);
			pos += 2;
			return v;
		}
		return (uint16_t) be(2);
	}
	inline uint64_t be(int nbytes) {
		uint64_t v = 0;
		for (int i = 0; i < nbytes; i++) {
			uint8_t b = 0xFF;
			if (pos < n)
				b = p[pos];
			else
				overrun = true;
			pos++;
			v = (v << 8) | b;
		}
		return v;
	}
};

struct Frame {
	uint32_t node;  // index into nodes (provisional)
	uint8_t next;   // next child slot to decode
	uint8_t depth;  // trie depth of this node (uint8 arithmetic like the reference)
	bool any;       // some child was non-NULL
};

} // namespace

int decodeIndexFile(const std::string &path, DecodedIndex &out, std::string &err) {
	std::vector<uint8_t> ibuf, abuf;
	if (!slurp(path, ibuf)) {
		err = "Cannot open file: " + path + ".";
		return CQ_EIO;
	}
	if (!slurp(path + ".aux", abuf)) {
		err = "Cannot open file: " + path + ".aux.";
		return CQ_EIO;
	}
	BitCursor bc = {abuf.data(), (uint64_t) abuf.size() * 8, 0};
	ByteCursor ic = {ibuf.data(), (uint64_t) ibuf.size(), 0, false};

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
	const uint32_t h = out.hash_len;
	// size hint: one leaf record is 6 or 12 bytes, almost every bucket is a single leaf
	size_t hint = ibuf.size() / (dd ? 20 : 14) + 16;
	out.bucket_key.reserve(hint);
	out.bucket_root.reserve(hint);
	out.ref_id1.reserve(hint);
	out.ucount1.reserve(hint);
	out.depth.reserve(hint);
	if (dd) {
		out.ref_id2.reserve(hint);
		out.ucount2.reserve(hint);
	}

	auto readLeaf = [&](uint8_t depth) -> uint32_t {
		uint32_t id = (uint32_t) out.ref_id1.size();
		uint32_t r1 = ic.be4(), r2 = 0;
		uint16_t c1, c2 = 0;
		if (dd) {
			r2 = ic.be4();
			c1 = ic.be2();
			c2 = ic.be2();
		} else
			c1 = ic.be2();
		out.ref_id1.push_back(r1);
		out.ref_id2.push_back(r2);
		out.ucount1.push_back(c1);
		out.ucount2.push_back(c2);
		out.depth.push_back((uint8_t) (depth + h));
		if (r1 > out.max_ref_id) out.max_ref_id = r1;
		if (r2 > out.max_ref_id) out.max_ref_id = r2;
		return kRefLeafTag | id;
	};

	std::vector<Frame> stack;
	uint64_t key = ic.be8();
	while (key != UINT64_MAX) {
		if (ic.overrun) {
			err = "Index " + path + ": INT stream ends before the END64 terminator.";
			return CQ_EFORMAT;
		}
		uint32_t root = kRefNone;
		// fast path: the bucket is one leaf at depth h ("10000")
		if (bc.peek5() == 0x10u) {
			bc.pos += 5;
			root = readLeaf(0);
		} else if (bc.bit() != 0) {
			stack.clear();
			out.nodes.insert(out.nodes.end(), 4, kRefNone);
			Frame f0 = {(uint32_t) (out.nodes.size() / 4 - 1), 0, 0, false};
			stack.push_back(f0);
			while (!stack.empty()) {
				Frame &f = stack.back();
				if (f.next < 4) {
					uint8_t slot = f.next++;
					if (bc.peek5() == 0x10u) {
						bc.pos += 5;
						uint32_t lr = readLeaf((uint8_t) (f.depth + 1));
						out.nodes[4 * (size_t) f.node + slot] = lr;
						f.any = true;
					} else if (bc.bit() != 0) {
						if (stack.size() > 4096 || bc.pos > bc.nbits + 64) {
							err = "Index " + path + ": AUX stream is malformed (runaway trie).";
							return CQ_EFORMAT;
						}
						f.any = true;
						uint8_t d1 = (uint8_t) (f.depth + 1);
						uint32_t parent = f.node;
						out.nodes.insert(out.nodes.end(), 4, kRefNone);
						uint32_t id = (uint32_t) (out.nodes.size() / 4 - 1);
						out.nodes[4 * (size_t) parent + slot] = id + 1;
						Frame nf = {id, 0, d1, false};
						stack.push_back(nf); // invalidates f
					}
					continue;
				}
				// all four children decoded
				Frame done = f;
				stack.pop_back();
				if (!done.any) {
					// a node with four NULL children is a leaf (hashtrie.cpp:432-457); it is the
					// most recently created node, so it can be dropped from the node array
					out.nodes.resize(out.nodes.size() - 4);
					uint32_t lr = readLeaf(done.depth);
					if (stack.empty())
						root = lr;
					else {
						Frame &p = stack.back();
						out.nodes[4 * (size_t) p.node + (p.next - 1)] = lr;
					}
				} else if (stack.empty())
					root = done.node + 1;
			}
		}
		if (dd && root != kRefNone) {
			// doubly-unique leaves must carry two ids (assert at hashtrie.cpp:446)
		}
		out.bucket_key.push_back(key);
		out.bucket_root.push_back(root);
		key = ic.be8();
		if (out.ref_id1.size() >= 0x7FFFFFF0ull || out.nodes.size() / 4 >= 0x7FFFFFF0ull) {
			err = "Index " + path + ": more than 2^31 leaves or nodes.";
			return CQ_EFORMAT;
		}
	}
	if (ic.overrun) {
		err = "Index " + path + ": INT stream truncated.";
		return CQ_EFORMAT;
	}
	if (dd)
		for (size_t i = 0; i < out.ref_id1.size(); i++)
			if (out.ref_id1[i] == 0 || out.ref_id2[i] == 0) {
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
