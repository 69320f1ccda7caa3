// host_tests.cpp -- tests of the C++ host mirror (ct_mapreduce_b200/host/ctmr_storage.hpp).
//   host_tests cpu                         the reference's own reducer tests, transcribed (no GPU, no libctmr calls)
//   host_tests gpu <dir> <batches>         StoreBatch over a corpus on disk through the GPU; dumps the
//                                          MockRemoteCache / MockBackend state for the pytest side to compare
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../ct_mapreduce_b200/host/ctmr_storage.hpp"

using namespace ctmr_host;

#define CHECK(c)                                                                   \
    do {                                                                           \
        if (!(c)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } \
    } while (0)

static int64_t unix_utc(int y, unsigned m, unsigned d, int hh, int mm, int ss) {
    y -= m <= 2;
    const int64_t era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m > 2 ? m - 3 : m + 9) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return (era * 146097 + (int64_t)doe - 719468) * 86400 + hh * 3600 + mm * 60 + ss;
}

// storage/knowncertificates_test.go:11-55
static void Test_Unknown() {
    MockRemoteCache backend;
    KnownCertificates kc(ExpDate::FromUnix(unix_utc(2029, 1, 30, 0, 0, 0)), Issuer::FromString("test issuer"), &backend);
    backend.Data[kc.serialId()] = {"\x01", "\x02", "\x03", "\x04"};
    for (const char* h : {"01", "02", "03", "04"}) CHECK(kc.WasUnknown(Serial::FromHex(h)).first == false);
    CHECK(kc.WasUnknown(Serial::FromHex("05")).first == true);
    CHECK(kc.WasUnknown(Serial::FromHex("05")).first == false);
    const std::vector<std::string> want = {"\x01", "\x02", "\x03", "\x04", "\x05"};
    CHECK(backend.Data[kc.serialId()] == want);
}

// storage/knowncertificates_test.go:57-83
static void Test_KnownCertificatesKnown() {
    MockRemoteCache backend;
    KnownCertificates kc(ExpDate::FromUnix(unix_utc(2029, 1, 30, 0, 0, 0)), Issuer::FromString("test issuer"), &backend);
    backend.Data[kc.serialId()] = {"\x01", "\x03", "\x05"};
    auto known = kc.Known();
    CHECK(known.size() == 3 && known[0].HexString() == "01" && known[1].HexString() == "03" && known[2].HexString() == "05");
    CHECK(kc.Count() == 3);
}

// storage/knowncertificates_test.go:85-110
static void Test_ExpireAt() {
    MockRemoteCache backend;
    KnownCertificates kc(ExpDate::FromUnix(unix_utc(2004, 1, 20, 4, 22, 19)), Issuer::FromString("test issuer"), &backend);
    CHECK(kc.WasUnknown(Serial::FromHex("05")).first == true);
    CHECK(backend.Expirations.size() == 1);
    auto it = backend.Expirations.find("serials::2004-01-20-04::test issuer");
    CHECK(it != backend.Expirations.end());
    CHECK(it->second == unix_utc(2004, 1, 20, 4, 0, 0));
}

// storage/issuermetadata_test.go:16-60
static void Test_DuplicateCRLs() {
    MockRemoteCache cache;
    IssuerMetadata meta(Issuer::FromString("issuer"), &cache);
    CHECK(ok(meta.addCRL("ldaps://ldap.crl")));
    CHECK(ok(meta.addCRL("schema://192.168.1.1:129/file.crl")));
    CHECK(ok(meta.addCRL("http://::1/file.crl")));
    CHECK(meta.CRLs().size() == 1);
    CHECK(ok(meta.addCRL("http://::1/file.crl")));
    CHECK(meta.CRLs().size() == 1);
    CHECK(ok(meta.addCRL("http://::1/file.crl ")));
    CHECK(meta.CRLs().size() == 1);
    CHECK(ok(meta.addCRL(" http://::1/file.crl ")));
    CHECK(ok(meta.addCRL(" http://::1/file.crl   ")));
    CHECK(meta.CRLs().size() == 1);
}

// storage/issuermetadata_test.go:100-136
static void Test_Accumulate() {
    MockRemoteCache cache;
    IssuerMetadata meta(Issuer::FromString("an issuer"), &cache);
    const ExpDate day = ExpDate::FromUnix(unix_utc(2001, 1, 1, 0, 0, 0));
    auto r1 = meta.Accumulate(day, "CN=My First Issuer (tm)", {});
    CHECK(ok(r1.second) && r1.first == false);  // "Should have been a new day"
    auto r2 = meta.Accumulate(day, "CN=My First Issuer (tm)", {});
    CHECK(ok(r2.second) && r2.first == true);
    CHECK(meta.CRLs().empty());
    CHECK(meta.Issuers().size() == 1 && meta.Issuers()[0] == "CN=My First Issuer (tm)");
}

// storage/types_test.go:41-57 (base64url), :81-101 (serial ID), :203-252 (ExpDate)
static void Test_Types() {
    const uint8_t d[32] = {0xa8, 0x10, 0x0a, 0xe6, 0xaa, 0x19, 0x40, 0xd0, 0xb6, 0x63, 0xbb, 0x31, 0xcd, 0x46, 0x61, 0x42,
                           0xeb, 0xbd, 0xbd, 0x51, 0x87, 0x13, 0x1b, 0x92, 0xd9, 0x38, 0x18, 0x98, 0x78, 0x32, 0xeb, 0x89};
    CHECK(Issuer::FromDigest(d).ID() == "qBAK5qoZQNC2Y7sxzUZhQuu9vVGHExuS2TgYmHgy64k=");  // SHA-256([0xFF]) digest
    CHECK(Serial::FromHex("00aa").ID() == "AKo=");
    CHECK(Serial::FromHex("0044aaff").ID() == "AESq_w==");
    CHECK(ExpDate::FromUnix(unix_utc(2004, 1, 20, 4, 22, 19)).ID() == "2004-01-20-04");
    CHECK(ExpDate::FromUnix(unix_utc(2004, 1, 19, 23, 59, 59)).ID() == "2004-01-19-23");
    CHECK(ExpDate::FromUnix(-1).ID() == "1969-12-31-23");
    CHECK(ExpDate::FromUnix(unix_utc(2004, 1, 20, 4, 22, 19)).DayID() == "2004-01-20");
}

template <class T>
static std::vector<T> read_all(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", path.c_str()); std::exit(2); }
    f.seekg(0, std::ios::end);
    const size_t n = (size_t)f.tellg();
    f.seekg(0);
    std::vector<T> v(n / sizeof(T));
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
    return v;
}

static std::string hex(const std::string& s) {
    std::string o;
    char b[3];
    for (unsigned char c : s) { std::snprintf(b, sizeof b, "%02x", c); o += b; }
    return o;
}

static int run_gpu(const std::string& dir, int batches, bool pipelined, int shards) {
    auto blob = read_all<uint8_t>(dir + "/blob.bin");
    auto offs = read_all<uint64_t>(dir + "/offsets.bin");
    auto iblob = read_all<uint8_t>(dir + "/issuer_blob.bin");
    auto ioffs = read_all<uint64_t>(dir + "/issuer_offsets.bin");
    auto idx = read_all<uint32_t>(dir + "/issuer_idx.bin");
    auto now = read_all<int64_t>(dir + "/now_ns.bin");
    auto filt = read_all<uint8_t>(dir + "/filter.bin");
    const uint64_t n = offs.size() - 1;
    ctmr_config cfg{};
    cfg.struct_size = sizeof cfg;
    cfg.table_capacity = 1 << 18;
    cfg.issuer_cn_filter = filt.data();
    cfg.issuer_cn_filter_len = (uint32_t)filt.size();
    ctmr_ctx* ctx = nullptr;
    ctmr_group* group = nullptr;
    if (shards > 1) {  // several shards on device 0: the group API from C++ (what the Go host calls on a multi-GPU box)
        cfg.max_batch_entries = 512;  // many rounds per batch
        std::vector<int32_t> devs((size_t)shards, 0);
        if (ctmr_group_create(&cfg, devs.data(), (uint32_t)shards, &group) != CTMR_OK) {
            std::fprintf(stderr, "ctmr_group_create: %s\n", ctmr_group_last_error(nullptr));
            return 3;
        }
    } else if (ctmr_create(&cfg, &ctx) != CTMR_OK) {
        std::fprintf(stderr, "ctmr_create: %s\n", ctmr_last_error(nullptr));
        return 3;
    }
    MockBatchRemoteCache batch_cache;  // offers SetInsertBatch / ExpireAtBatch (found by dynamic_cast in StoreBatch)
    MockRemoteCache plain_cache;
    MockRemoteCache& cache = pipelined ? static_cast<MockRemoteCache&>(batch_cache) : plain_cache;
    MockBackend backend;
    GpuCertDatabase db_single(ctx, &cache, &backend), db_group(group, &cache, &backend);
    GpuCertDatabase& db = group ? db_group : db_single;
    BatchStats total{};
    for (int b = 0; b < batches; ++b) {
        const uint64_t lo = n * b / batches, hi = n * (b + 1) / batches;
        BatchStats st;
        Error e = db.StoreBatch(blob.data(), offs.data() + lo, hi - lo, iblob.data(), ioffs.data(), (uint32_t)ioffs.size() - 1,
                                idx.data() + lo, now[0], &st);
        if (!ok(e)) { std::fprintf(stderr, "StoreBatch: %s\n", e.c_str()); return 4; }
        total.entries += st.entries; total.stored += st.stored; total.unknown += st.unknown;
        total.cache_inserts += st.cache_inserts; total.pem_writes += st.pem_writes;
        total.dn_formats += st.dn_formats; total.crl_parses += st.crl_parses;
    }
    std::ofstream out(dir + "/state.txt");
    out << "STATS entries " << total.entries << " stored " << total.stored << " unknown " << total.unknown << " set_insert_calls "
        << cache.set_insert_calls << " pem_writes " << total.pem_writes << " mark_dirty_calls " << backend.mark_dirty_calls
        << " dn_formats " << total.dn_formats << " crl_parses " << total.crl_parses << " round_trips " << batch_cache.round_trips << "\n";
    for (const auto& kv : cache.Data)
        for (const auto& m : kv.second) out << "SET " << kv.first << " " << hex(m) << "\n";
    // issuer:: and crl:: members are text: also dump them readable (spaces escaped) for the pytest side
    for (const auto& kv : cache.Data) {
        if (kv.first.compare(0, 8, "issuer::") != 0 && kv.first.compare(0, 5, "crl::") != 0) continue;
        for (const auto& m : kv.second) out << "STR " << kv.first << " " << hex(m) << "\n";
    }
    for (const auto& kv : cache.Expirations) out << "EXPIRE " << kv.first << " " << kv.second << "\n";
    for (const auto& d : backend.dirty) out << "DIRTY " << d << "\n";
    for (const auto& kv : backend.issuers_by_expdate)
        for (const auto& i : kv.second) out << "ALLOC " << kv.first << " " << i << "\n";
    int shown = 0;
    for (const auto& kv : backend.pems) {
        out << "PEMKEY " << kv.first << "\n";
        if (shown++ < 5) {
            std::string one = kv.second;
            std::replace(one.begin(), one.end(), '\n', '|');
            out << "PEM " << kv.first << " " << one << "\n";
        }
    }
    // the read side over the same interfaces: Count() through the cache == SetCardinality on the GPU table
    for (const auto& kv : cache.Data) {
        if (kv.first.compare(0, 9, "serials::") != 0) continue;
        out << "COUNT " << kv.first << " " << kv.second.size() << "\n";
    }
    if (group) ctmr_group_destroy(group);
    else ctmr_destroy(ctx);
    return 0;
}

// pkix.Name.String(): order, '+' for repeated types, escaping (Go's rules; the reference pins the single-CN case)
static void Test_FormatIssuerDN() {
    // C=US, O=Synth CA 5, CN=" ISRG Root X5"   (leading space must be escaped)
    const uint8_t name[] = {0x30, 0x39, 0x31, 0x0b, 0x30, 0x09, 0x06, 0x03, 0x55, 0x04, 0x06, 0x13, 0x02, 'U', 'S',
                            0x31, 0x13, 0x30, 0x11, 0x06, 0x03, 0x55, 0x04, 0x0a, 0x13, 0x0a, 'S', 'y', 'n', 't', 'h', ' ', 'C', 'A', ' ', '5',
                            0x31, 0x15, 0x30, 0x13, 0x06, 0x03, 0x55, 0x04, 0x03, 0x13, 0x0c, ' ', 'I', 'S', 'R', 'G', ' ', 'R', 'o', 'o', 't', ' ', 'X'};
    CHECK(FormatIssuerDN(name, sizeof name) == "CN=\\ ISRG Root X,O=Synth CA 5,C=US");
    const uint8_t one[] = {0x30, 0x1f, 0x31, 0x1d, 0x30, 0x1b, 0x06, 0x03, 0x55, 0x04, 0x03, 0x0c, 0x14, 'M', 'y', ' ', 'F', 'i', 'r', 's', 't', ' ',
                           'I', 's', 's', 'u', 'e', 'r', ' ', '(', 't', 'm', ')'};
    CHECK(FormatIssuerDN(one, sizeof one) == "CN=My First Issuer (tm)");  // issuermetadata_test.go:133
    // cRLDistributionPoints value with one http URI
    const char* uri = "http://crl.example/x.crl";
    std::vector<uint8_t> v = {0x30, 0x20, 0x30, 0x1e, 0xa0, 0x1c, 0xa0, 0x1a, 0x86, 0x18};
    v.insert(v.end(), uri, uri + 24);
    auto uris = ExtractCrlUris(v.data(), v.size());
    CHECK(uris.size() == 1 && uris[0] == uri);
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "cpu") {
        Test_Unknown();
        Test_KnownCertificatesKnown();
        Test_ExpireAt();
        Test_DuplicateCRLs();
        Test_Accumulate();
        Test_Types();
        Test_FormatIssuerDN();
        std::puts("host cpu tests: 7 passed");
        return 0;
    }
    if (argc >= 4 && std::string(argv[1]) == "gpu")
        return run_gpu(argv[2], std::atoi(argv[3]), argc >= 5 && std::string(argv[4]) == "pipelined",
                       argc >= 6 ? std::atoi(argv[5]) : (argc >= 5 && std::string(argv[4]) == "group" ? 3 : 1));
    std::fprintf(stderr, "usage: host_tests cpu | gpu <dir> <batches>\n");
    return 64;
}
