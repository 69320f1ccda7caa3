// cgo binding of include/ctmr_frontend.h: the CT wire-format front end (SURVEY.md §8(f)-2).
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain.  This is what a maintainer of
// jcjones/ct-mapreduce adds next to go/ctmr/ctmr.go to let the downloader hand get-entries bodies to the GPU
// instead of running ct.LogEntryFromLeaf per entry (cmd/ct-fetch/ct-fetch.go:446-484).
package ctmr

/*
#include "ctmr_frontend.h"
*/
import "C"

import (
	"bytes"
	"fmt"
	"unsafe"
)

// EntryStatus mirrors CTMR_FE_*.
type EntryStatus uint8

const (
	FeOK EntryStatus = iota
	FeBadBase64
	FeBadLeaf
	FeUnknownType
	FeBadExtra
	FeBadCert
)

// RawPages accumulates get-entries response bodies in pinned memory (ctmr_host_alloc) together with the
// spans of their leaf_input / extra_data strings.  The downloader appends each body as it arrives
// (after LogClient's HTTP GET, before any JSON decoding) and flushes with Ctx.ProcessRaw.
type RawPages struct {
	Text      unsafe.Pointer // pinned, Cap bytes
	Cap, Used uint64
	LeafOff   []uint64
	LeafLen   []uint32
	ExtraOff  []uint64
	ExtraLen  []uint32
}

var (
	leafKey  = []byte(`"leaf_input":"`)
	extraKey = []byte(`"extra_data":"`)
)

// Append copies one response body into the pinned text and records its string spans.  A body whose
// strings carry JSON escapes (no CT log emits any) must be unescaped by the caller first.
func (p *RawPages) Append(body []byte) error {
	if p.Used+uint64(len(body)) > p.Cap {
		return fmt.Errorf("raw page buffer full")
	}
	dst := unsafe.Slice((*byte)(unsafe.Add(p.Text, p.Used)), len(body))
	copy(dst, body)
	scan := func(key []byte, offs *[]uint64, lens *[]uint32) error {
		for at := 0; ; {
			i := bytes.Index(body[at:], key)
			if i < 0 {
				return nil
			}
			start := at + i + len(key)
			end := bytes.IndexByte(body[start:], '"')
			if end < 0 || bytes.IndexByte(body[start:start+end], '\\') >= 0 {
				return fmt.Errorf("malformed or escaped base64 string at %d", start)
			}
			*offs = append(*offs, p.Used+uint64(start))
			*lens = append(*lens, uint32(end))
			at = start + end
		}
	}
	if err := scan(leafKey, &p.LeafOff, &p.LeafLen); err != nil {
		return err
	}
	if err := scan(extraKey, &p.ExtraOff, &p.ExtraLen); err != nil {
		return err
	}
	if len(p.LeafOff) != len(p.ExtraOff) {
		return fmt.Errorf("get-entries body with %d leaf_input and %d extra_data strings", len(p.LeafOff), len(p.ExtraOff))
	}
	p.Used += uint64(len(body))
	return nil
}

// RawResult holds the per-entry outputs of ProcessRaw (ctmr_raw_out).
type RawResult struct {
	Result                  // the path's outputs, as for ProcessBatch
	EntryStatus, EntryType  []uint8
	TimestampMs             []uint64
	Issuer                  []uint32
	LeafSrc                 []uint8
	LeafOff, LeafLen        []uint32
}

// ProcessRaw = GetRawEntries' base64 decode + ct.LogEntryFromLeaf + insertCTWorker + Store decisions for
// every entry of the accumulated pages, in page order.
func (c *Ctx) ProcessRaw(p *RawPages, nowUnixNs int64, r *RawResult) error {
	n := len(p.LeafOff)
	if n == 0 {
		return nil
	}
	var b C.ctmr_raw_batch
	b.text = (*C.uint8_t)(p.Text)
	b.text_bytes = C.uint64_t(p.Used)
	b.leaf_input_off = (*C.uint64_t)(unsafe.Pointer(&p.LeafOff[0]))
	b.leaf_input_len = (*C.uint32_t)(unsafe.Pointer(&p.LeafLen[0]))
	b.extra_data_off = (*C.uint64_t)(unsafe.Pointer(&p.ExtraOff[0]))
	b.extra_data_len = (*C.uint32_t)(unsafe.Pointer(&p.ExtraLen[0]))
	b.n = C.uint64_t(n)
	b.now_unix_ns = C.int64_t(nowUnixNs)
	var o C.ctmr_raw_out
	o.path = r.Result.cOut() // the same pointer table ProcessBatch fills (ctmr.go)
	o.entry_status = (*C.uint8_t)(unsafe.Pointer(&r.EntryStatus[0]))
	o.entry_type = (*C.uint8_t)(unsafe.Pointer(&r.EntryType[0]))
	o.timestamp_ms = (*C.uint64_t)(unsafe.Pointer(&r.TimestampMs[0]))
	o.issuer = (*C.uint32_t)(unsafe.Pointer(&r.Issuer[0]))
	o.leaf_src = (*C.uint8_t)(unsafe.Pointer(&r.LeafSrc[0]))
	o.leaf_off = (*C.uint32_t)(unsafe.Pointer(&r.LeafOff[0]))
	o.leaf_len = (*C.uint32_t)(unsafe.Pointer(&r.LeafLen[0]))
	return c.err(C.ctmr_process_raw(c.h, &b, &o), "ctmr_process_raw")
}
