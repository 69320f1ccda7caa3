package ctmr

// GPUDatabase implements storage.CertDatabase (storage/types.go:70-81) -- all nine methods -- and adds StoreBatch.
// It keeps storage.CertDatabase, storage.RemoteCache and storage.StorageBackend byte-identical and only replaces
// WHO decides "parse ok? filtered? was unknown? first (issuer, hour)? string seen before?" -- the GPU(s) -- while
// every side effect still goes through the reference's own interfaces, in entry order:
//
//	Store                      per-entry fallback = the stock FilesystemDatabase.Store (used for the entries the
//	                           GPU path declines: serials longer than CTMR_MAX_SERIAL octets, which the reference
//	                           accepts, storage/types.go:171-178)
//	StoreBatch                 the accelerated path: one ctmr(_group)_process_batch call per drained batch
//	the other eight methods    delegated unchanged to the wrapped FilesystemDatabase
//
// NOT COMPILED IN THIS REPOSITORY (no Go toolchain in the build image; `go vet` / `go build` run where Go exists).
// Wiring in cmd/ct-fetch (replaces StartDatabaseThreads' N x insertCTWorker, ct-fetch.go:140-145):
//
//	func (ld *LogSyncEngine) insertCTBatcher(gpu *ctmr.GPUDatabase) {
//	    for {
//	        batch := ld.drain(ld.entryChan, 16384, 50*time.Millisecond) // ct-fetch.go:132: the channel depth is one batch
//	        if batch == nil { return }
//	        if err := gpu.StoreBatch(batch, time.Now()); err != nil { glog.Errorf("StoreBatch: %v", err) }
//	    }
//	}

import (
	"context"
	"encoding/base64"
	"fmt"
	"net/url"
	"time"

	"github.com/google/certificate-transparency-go/x509"
	"github.com/jcjones/ct-mapreduce/storage"
)

type Entry struct {
	Cert      *x509.Certificate // nil for precerts not yet parsed: LeafDER is what counts
	LeafDER   []byte            // X509Cert.Raw or Precert.Submitted.Data (ct-fetch.go:198-204)
	IssuerDER []byte            // Chain[0].Data, nil when len(Chain) < 1 (ct-fetch.go:215-219)
	LogURL    string
	Index     int64
}

// BatchRemoteCache is an OPTIONAL extension a RemoteCache may implement (discovered by type assertion, so
// storage.RemoteCache itself stays as it is): one pipelined round trip per batch instead of one per new certificate.
type BatchRemoteCache interface {
	SetInsertBatch(keys []string, entries []string) ([]bool, error)
	ExpireAtBatch(keys []string, at []time.Time) error
}

type GPUDatabase struct {
	inner   storage.CertDatabase // the stock FilesystemDatabase: fallback Store + the eight untouched methods
	db      *DB
	cache   storage.RemoteCache
	backend storage.StorageBackend
}

func NewGPUDatabase(inner storage.CertDatabase, cache storage.RemoteCache, backend storage.StorageBackend, db *DB) *GPUDatabase {
	return &GPUDatabase{inner: inner, db: db, cache: cache, backend: backend}
}

// ---- storage.CertDatabase: delegated ---------------------------------------------------------------------
func (g *GPUDatabase) Cleanup() error                                   { return g.inner.Cleanup() }
func (g *GPUDatabase) SaveLogState(l *storage.CertificateLog) error      { return g.inner.SaveLogState(l) }
func (g *GPUDatabase) GetLogState(u *url.URL) (*storage.CertificateLog, error) { return g.inner.GetLogState(u) }
func (g *GPUDatabase) ListExpirationDates(nb time.Time) ([]storage.ExpDate, error) {
	return g.inner.ListExpirationDates(nb)
}
func (g *GPUDatabase) ListIssuersForExpirationDate(e storage.ExpDate) ([]storage.Issuer, error) {
	return g.inner.ListIssuersForExpirationDate(e)
}
func (g *GPUDatabase) GetKnownCertificates(e storage.ExpDate, i storage.Issuer) *storage.KnownCertificates {
	return g.inner.GetKnownCertificates(e, i) // reads Redis, which StoreBatch keeps exact (one SADD per new certificate)
}
func (g *GPUDatabase) GetIssuerMetadata(i storage.Issuer) *storage.IssuerMetadata { return g.inner.GetIssuerMetadata(i) }
func (g *GPUDatabase) GetIssuerAndDatesFromCache() ([]storage.IssuerDate, error) {
	return g.inner.GetIssuerAndDatesFromCache()
}

// Store is the reference's own per-entry path.  Entries stored this way are NOT in the GPU tables, which is safe:
// Redis stays the source of truth (SetInsert answers for them), and the only entries routed here are the ones the
// GPU declined, which it will decline again.
func (g *GPUDatabase) Store(c *x509.Certificate, issuer *x509.Certificate, logURL string, id int64) error {
	return g.inner.Store(c, issuer, logURL, id)
}

// ---- the accelerated path ------------------------------------------------------------------------------------
// StoreBatch reproduces, for a whole batch, exactly the calls FilesystemDatabase.Store makes
// (filesystemdatabase.go:158-211) for the entries the GPU reports as reaching Store:
//   - was_unknown        -> cache.SetInsert(serials::<exp>::<issuer>, serial)   (knowncertificates.go:39)
//   - first (exp,issuer) -> cache.ExpireAt(key, expDate) + backend.AllocateExpDateAndIssuer (knowncertificates.go:44-47, filesystemdatabase.go:189-195)
//   - first DN / CRL-DP / (exp,issuer) -> IssuerMetadata.Accumulate(cert) on the reference object: the only entries
//     for which it has any effect, all others are memo hits there (issuermetadata.go:92-138)
//   - was_unknown        -> backend.StoreCertificatePEM with the PEM text the GPU encoded (filesystemdatabase.go:197-201)
//   - every stored entry -> backend.MarkDirty(YYYY-MM-DD), once per distinct day    (filesystemdatabase.go:205)
func (g *GPUDatabase) StoreBatch(entries []Entry, now time.Time) error {
	b, derBytes, issuerIdxOf := pack(entries) // leaf DERs -> pinned blob, distinct Chain[0] DERs -> issuer table
	defer b.Free()
	r := NewResult(b.N, derBytes)
	defer r.Free()
	dense, err := g.db.RegisterIssuers(b.IssuerBlob, b.IssuerOffs) // memoised in the library: GPU work only for new issuers
	if err != nil {
		return err
	}
	if err := g.db.ProcessBatch(b, now.UnixNano(), r); err != nil {
		return err // batch-level failure (CUDA, table full): same severity as a Redis outage
	}
	issuerID := map[uint32]string{} // Issuer.ID() = base64url(SHA-256(SPKI)): the digest comes back from the library
	idOf := func(k uint32) (string, error) {
		if s, ok := issuerID[k]; ok {
			return s, nil
		}
		dg, err := g.db.IssuerDigest(dense[k])
		if err != nil {
			return "", err
		}
		issuerID[k] = base64.URLEncoding.EncodeToString(dg[:])
		return issuerID[k], nil
	}
	var keys, members []string
	var newIdx []int
	dirty := map[string]struct{}{}
	for i, e := range entries {
		switch Status(r.Status[i]) {
		case StOK:
		case StSerialTooLong: // the reference has no such limit: stock per-entry path
			if cert, perr := parsed(e); perr == nil {
				if icert, ierr := x509.ParseCertificate(e.IssuerDER); ierr == nil {
					if err := g.inner.Store(cert, icert, e.LogURL, e.Index); err != nil {
						return err
					}
				}
			}
			continue
		default:
			continue // logged and counted on the Go side exactly as ct-fetch.go:206-232 does, from r.Status[i]
		}
		expTime := time.Unix(r.ExpHour[i]*3600, 0).UTC()
		dirty[expTime.Format("2006-01-02")] = struct{}{}
		if r.WasUnknown[i] == 0 {
			continue
		}
		id, err := idOf(issuerIdxOf[i])
		if err != nil {
			return err
		}
		expDate := storage.NewExpDateFromTime(expTime)
		serial := storage.NewSerialFromBytes(e.LeafDER[r.SerialOff[i] : r.SerialOff[i]+r.SerialLen[i]])
		keys = append(keys, "serials::"+expDate.ID()+"::"+id)
		members = append(members, serial.BinaryString())
		newIdx = append(newIdx, i)
	}
	// Redis: the new members, pipelined when the cache offers it
	if bc, ok := g.cache.(BatchRemoteCache); ok {
		if _, err := bc.SetInsertBatch(keys, members); err != nil {
			return err
		}
	} else {
		for j := range keys {
			if _, err := g.cache.SetInsert(keys[j], members[j]); err != nil {
				return err
			}
		}
	}
	ctx := context.Background()
	for j, i := range newIdx {
		e := entries[i]
		id, _ := idOf(issuerIdxOf[i])
		issuer := storage.NewIssuerFromString(id)
		expDate := storage.NewExpDateFromTime(time.Unix(r.ExpHour[i]*3600, 0).UTC())
		if r.FirstIssuerHour[i] == 1 {
			if err := g.cache.ExpireAt(keys[j], expDate.ExpireTime()); err != nil {
				return err
			}
			if err := g.backend.AllocateExpDateAndIssuer(ctx, expDate, issuer); err != nil {
				return err
			}
		}
		if r.FirstIssuerHour[i] == 1 || r.FirstIssuerDN[i] == 1 || r.FirstCrlDp[i] == 1 {
			cert, perr := parsed(e) // O(issuers x few) parses per run instead of one per new certificate
			if perr != nil {
				return fmt.Errorf("entry %d: GPU accepted a certificate ct-go rejects: %v", e.Index, perr)
			}
			if _, err := g.inner.GetIssuerMetadata(issuer).Accumulate(cert); err != nil {
				return err
			}
		}
		serial := storage.NewSerialFromBytes(e.LeafDER[r.SerialOff[i] : r.SerialOff[i]+r.SerialLen[i]])
		if err := g.backend.StoreCertificatePEM(ctx, serial, expDate, issuer, r.PEMOf(i)); err != nil {
			return err
		}
	}
	for day := range dirty {
		if err := g.backend.MarkDirty(day); err != nil {
			return err
		}
	}
	return nil
}

func parsed(e Entry) (*x509.Certificate, error) {
	if e.Cert != nil {
		return e.Cert, nil
	}
	return x509.ParseCertificate(e.LeafDER)
}

// Free releases the pinned buffers of a packed batch.
func (b *Batch) Free() {
	HostFree(b.Blob)
	HostFree(b.Offsets)
	HostFree(b.IssuerIdx)
}
