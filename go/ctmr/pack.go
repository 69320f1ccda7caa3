package ctmr

import "unsafe"

// pack lays a drained batch out the way ctmr_process_batch takes it: leaf DERs back to back in pinned memory,
// uint64 offsets, and per entry an index into the batch's table of DISTINCT Chain[0] certificates (a CT log has a
// handful of issuers; the library parses and SPKI-hashes each distinct one once, ever).
func pack(entries []Entry) (b *Batch, derBytes uint64, issuerIdxOf []uint32) {
	n := len(entries)
	for _, e := range entries {
		derBytes += uint64(len(e.LeafDER))
	}
	b = &Batch{N: n, Blob: HostAlloc(int(derBytes) + 64), Offsets: HostAlloc(8 * (n + 1)), IssuerIdx: HostAlloc(4 * n)}
	blob := unsafe.Slice((*byte)(b.Blob), int(derBytes))
	offs := unsafe.Slice((*uint64)(b.Offsets), n+1)
	idx := unsafe.Slice((*uint32)(b.IssuerIdx), n)
	issuerIdxOf = make([]uint32, n)
	seen := map[string]uint32{}
	b.IssuerOffs = []uint64{0}
	var at uint64
	for i, e := range entries {
		offs[i] = at
		copy(blob[at:], e.LeafDER)
		at += uint64(len(e.LeafDER))
		if e.IssuerDER == nil {
			idx[i] = IssuerNone // len(Chain) < 1: the entry gets status NO_ISSUER, as ct-fetch.go:215-219 skips it
			continue
		}
		k, ok := seen[string(e.IssuerDER)]
		if !ok {
			k = uint32(len(seen))
			seen[string(e.IssuerDER)] = k
			b.IssuerBlob = append(b.IssuerBlob, e.IssuerDER...)
			b.IssuerOffs = append(b.IssuerOffs, uint64(len(b.IssuerBlob)))
		}
		idx[i] = k
		issuerIdxOf[i] = k
	}
	offs[n] = at
	return b, derBytes, issuerIdxOf
}
