// Package ctmr binds libctmr.so (include/ctmr.h) with cgo: the B200-native replacement for the
// loop body of insertCTWorker (cmd/ct-fetch/ct-fetch.go:191-245) and FilesystemDatabase.Store
// (storage/filesystemdatabase.go:158-211).
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain.  It is the binding a
// maintainer of jcjones/ct-mapreduce adds; see INTEGRATION.md.
package ctmr

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../ct_mapreduce_b200 -lctmr -Wl,-rpath,${SRCDIR}/../../ct_mapreduce_b200
#include <stdlib.h>
#include "ctmr.h"
*/
import "C"

import (
	"fmt"
	"unsafe"
)

// Status mirrors CTMR_ST_*.
type Status uint8

const (
	StOK Status = iota
	StParseErr
	StFilterCA
	StFilterExpired
	StFilterCN
	StNoIssuer
	StIssuerParseErr
	StSerialTooLong
)

const IssuerNone = 0xFFFFFFFF

// Ctx owns one GPU's known-certificate state.
type Ctx struct{ h *C.ctmr_ctx }

type Config struct {
	Device            int
	TableCapacity     uint64
	MaxBatchEntries   uint64
	MaxBatchBytes     uint64
	IssuerCNFilter    string // *ctconfig.IssuerCNFilter, passed verbatim (split on ',' without trimming)
	LogExpiredEntries bool   // *ctconfig.LogExpiredEntries
	NoFingerprint     bool
}

func New(c Config) (*Ctx, error) {
	var cfg C.ctmr_config
	cfg.struct_size = C.uint32_t(unsafe.Sizeof(cfg))
	cfg.device = C.int32_t(c.Device)
	cfg.table_capacity = C.uint64_t(c.TableCapacity)
	cfg.max_batch_entries = C.uint64_t(c.MaxBatchEntries)
	cfg.max_batch_bytes = C.uint64_t(c.MaxBatchBytes)
	var filt unsafe.Pointer
	if len(c.IssuerCNFilter) > 0 {
		filt = C.CBytes([]byte(c.IssuerCNFilter))
		defer C.free(filt)
		cfg.issuer_cn_filter = (*C.uint8_t)(filt)
		cfg.issuer_cn_filter_len = C.uint32_t(len(c.IssuerCNFilter))
	}
	if c.LogExpiredEntries {
		cfg.log_expired_entries = 1
	}
	if c.NoFingerprint {
		cfg.flags |= C.CTMR_F_NO_FINGERPRINT
	}
	var h *C.ctmr_ctx
	if rc := C.ctmr_create(&cfg, &h); rc != 0 {
		return nil, fmt.Errorf("ctmr_create: %d: %s", int(rc), C.GoString(C.ctmr_last_error(nil)))
	}
	return &Ctx{h: h}, nil
}

func (c *Ctx) Close() { C.ctmr_destroy(c.h); c.h = nil }

func (c *Ctx) err(rc C.int, what string) error {
	if rc == 0 {
		return nil
	}
	return fmt.Errorf("%s: %d: %s", what, int(rc), C.GoString(C.ctmr_last_error(c.h)))
}

// Batch is packed by the batcher goroutine into C (pinned) memory: cgo forbids C from keeping Go
// pointers, and pinned memory lets the H2D copy overlap the kernels.
type Batch struct {
	N          int
	Blob       unsafe.Pointer // ctmr_host_alloc'ed, leaf DERs back to back
	Offsets    unsafe.Pointer // uint64[N+1]
	IssuerIdx  unsafe.Pointer // uint32[N], index into the batch's distinct Chain[0] list, IssuerNone = no chain
	IssuerBlob []byte
	IssuerOffs []uint64
}

// Result arrays are Go-owned; the library only writes them during the call.
type Result struct {
	Status          []uint8
	SHA256          []byte // N*32
	ExpHour         []int64
	SerialOff       []uint32
	SerialLen       []uint32
	WasUnknown      []uint8
	FirstIssuerHour []uint8
}

func NewResult(n int) *Result {
	return &Result{make([]uint8, n), make([]byte, 32*n), make([]int64, n), make([]uint32, n), make([]uint32, n),
		make([]uint8, n), make([]uint8, n)}
}

// cOut is the ctmr_out pointer table over the Go-owned result arrays (shared with ProcessRaw, frontend.go).
func (r *Result) cOut() C.ctmr_out {
	return C.ctmr_out{
		status:            (*C.uint8_t)(unsafe.Pointer(&r.Status[0])),
		sha256:            (*C.uint8_t)(unsafe.Pointer(&r.SHA256[0])),
		exp_hour:          (*C.int64_t)(unsafe.Pointer(&r.ExpHour[0])),
		serial_off:        (*C.uint32_t)(unsafe.Pointer(&r.SerialOff[0])),
		serial_len:        (*C.uint32_t)(unsafe.Pointer(&r.SerialLen[0])),
		was_unknown:       (*C.uint8_t)(unsafe.Pointer(&r.WasUnknown[0])),
		first_issuer_hour: (*C.uint8_t)(unsafe.Pointer(&r.FirstIssuerHour[0])),
	}
}

// ProcessBatch = parse + certIsFilteredOut + Store decisions for every entry of the batch.
// nowUnixNs replaces time.Now() at ct-fetch.go:52.
func (c *Ctx) ProcessBatch(b *Batch, nowUnixNs int64, r *Result) error {
	if b.N == 0 {
		return nil
	}
	out := r.cOut()
	var ib *C.uint8_t
	var io *C.uint64_t
	if len(b.IssuerOffs) > 1 {
		ib = (*C.uint8_t)(unsafe.Pointer(&b.IssuerBlob[0]))
		io = (*C.uint64_t)(unsafe.Pointer(&b.IssuerOffs[0]))
	}
	rc := C.ctmr_process_batch(c.h, (*C.uint8_t)(b.Blob), (*C.uint64_t)(b.Offsets), C.uint64_t(b.N), ib, io,
		C.uint32_t(len(b.IssuerOffs)-1), (*C.uint32_t)(b.IssuerIdx), C.int64_t(nowUnixNs), &out)
	return c.err(rc, "ctmr_process_batch")
}

// IssuerCounts = per issuer, the sum over expDates of KnownCertificates.Count()
// (cmd/storage-statistics/storage-statistics.go:44-53).  Keys are raw SHA-256(SPKI) digests;
// base64.URLEncoding of a key is Issuer.ID().
func (c *Ctx) IssuerCounts() (map[[32]byte]uint64, error) {
	n := C.size_t(C.ctmr_issuer_count(c.h))
	if n == 0 {
		return map[[32]byte]uint64{}, nil
	}
	dig := make([]byte, 32*int(n))
	cnt := make([]uint64, int(n))
	rc := C.ctmr_issuer_counts(c.h, (*C.uint8_t)(unsafe.Pointer(&dig[0])), (*C.uint64_t)(unsafe.Pointer(&cnt[0])), &n)
	if err := c.err(rc, "ctmr_issuer_counts"); err != nil {
		return nil, err
	}
	out := make(map[[32]byte]uint64, int(n))
	for i := 0; i < int(n); i++ {
		var k [32]byte
		copy(k[:], dig[32*i:32*i+32])
		out[k] = cnt[i]
	}
	return out, nil
}

// SetCardinality = RemoteCache.SetCardinality("serials::<expDate>::<issuer>") answered from the GPU table.
func (c *Ctx) SetCardinality(expHour int64, issuerDigest [32]byte) (uint64, error) {
	var v C.uint64_t
	rc := C.ctmr_set_cardinality(c.h, C.int64_t(expHour), (*C.uint8_t)(unsafe.Pointer(&issuerDigest[0])), &v)
	return uint64(v), c.err(rc, "ctmr_set_cardinality")
}

func HostAlloc(n int) unsafe.Pointer { return C.ctmr_host_alloc(C.size_t(n)) }
func HostFree(p unsafe.Pointer)      { C.ctmr_host_free(p) }
