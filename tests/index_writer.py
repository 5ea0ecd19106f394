"""Format-exact CAMMiQ index writer for TESTS (reference writer
hashtrie.cpp:595-700, binaryio.cpp:3-134).  Lets tests build adversarial indices the
reference builder would never emit at small scale: dense keys, branching buckets, keys
present in both tables, deep tries."""

CODE = {65: 0, 67: 1, 71: 2, 84: 3}


def write_index(path, h, entries, doubly):
    """entries: iterable of (key: bytes of ACGT with len >= h, rid1, rid2, ucount1, ucount2).
    Buckets are emitted in first-appearance order; leaves in pre-order (A<C<G<T)."""
    buckets, order = {}, []
    for key, r1, r2, c1, c2 in entries:
        assert len(key) >= h
        bk = 0
        for ch in key[:h]:
            bk = (bk << 2) | CODE[ch]
        if bk not in buckets:
            buckets[bk] = {}
            order.append(bk)
        node = buckets[bk]
        for ch in key[h:]:
            assert "leaf" not in node, "keys must be prefix-free"
            node = node.setdefault(CODE[ch], {})
        assert not node, "keys must be prefix-free / distinct"
        node["leaf"] = (r1, r2, c1, c2)
    bits, ints = [], bytearray()
    bits += [1 if doubly else 0]
    bits += [(64 >> i) & 1 for i in range(6, -1, -1)]
    bits += [(h >> i) & 1 for i in range(7, -1, -1)]

    def emit(node):
        bits.append(1)
        if "leaf" in node:
            bits.extend([0, 0, 0, 0])
            r1, r2, c1, c2 = node["leaf"]
            ints.extend(r1.to_bytes(4, "big"))
            if doubly:
                ints.extend(r2.to_bytes(4, "big"))
                ints.extend(c1.to_bytes(2, "big"))
                ints.extend(c2.to_bytes(2, "big"))
            else:
                ints.extend(c1.to_bytes(2, "big"))
            return
        for c in range(4):
            if c in node:
                emit(node[c])
            else:
                bits.append(0)

    for bk in order:
        ints.extend(bk.to_bytes(8, "big"))
        emit(buckets[bk])
    bits += [1] * 72
    ints.extend(b"\xff" * 10)
    nbytes = len(bits) // 8  # the reference writer drops a trailing partial byte
    aux = bytearray(nbytes)
    for i in range(nbytes * 8):
        if bits[i]:
            aux[i >> 3] |= 0x80 >> (i & 7)
    with open(path, "wb") as f:
        f.write(ints)
    with open(path + ".aux", "wb") as f:
        f.write(aux)
