// seam_bench.cpp -- the reference's per-chunk benchmark loop (src/c++/benchmark.cpp:134-190: one timed call per chunk,
// bits/symbol and MiB/s averaged over the chunks, benchmark.cpp:269-295) driven through the two drop-in seams of
// the C++ functions of rnsc_signal_compression.cpp (pod5::/pgnano::compress_signal, decompress_signal) and
// the C API names of libpod5_rnsc.so (pod5_vbz_compress_signal, pod5_pinanoraw_compress_signal, pod5_vbz_decompress_signal).
// Input: a chunk file `[u32 n][int16 x n]...` (benchmark.cpp:26-49). Every decoded chunk is compared with its input.
//   seam_bench <chunkfile> [threads]       prints one JSON object
#include "pod5_format/signal_compression.h"
#include "pod5_format/pgnano/pgnano.h"
#include "pod5_rnsc.h"
#include "rnsc.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

struct Chunk { std::vector<int16_t> x; };
struct Acc { double bits = 0, mibs_c = 0, mibs_d = 0, t_c = 0, t_d = 0; size_t chunks = 0, samples = 0, bytes = 0; bool ok = true; std::string err; };

using Clock = std::chrono::steady_clock;
static double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

static std::vector<Chunk> load(const char* path)
{
    std::vector<Chunk> v;
    FILE* f = fopen(path, "rb");
    if (!f) return v;
    uint32_t n;
    while (fread(&n, 4, 1, f) == 1) { Chunk c; c.x.resize(n); if (n && fread(c.x.data(), 2, n, f) != n) break; v.push_back(std::move(c)); }
    fclose(f);
    return v;
}

template <typename Enc, typename Dec>
static Acc run(const std::vector<Chunk>& chunks, Enc enc, Dec dec)
{
    Acc a;
    std::vector<int16_t> back;
    for (const Chunk& c : chunks) {
        const size_t n = c.x.size();
        std::vector<uint8_t> comp;
        auto t0 = Clock::now();
        std::string e = enc(c.x, comp);
        auto t1 = Clock::now();
        if (!e.empty()) { a.ok = false; a.err = e; break; }
        back.assign(n, 0);
        auto t2 = Clock::now();
        e = dec(comp, back);
        auto t3 = Clock::now();
        if (!e.empty()) { a.ok = false; a.err = e; break; }
        if (n && memcmp(back.data(), c.x.data(), 2 * n) != 0) { a.ok = false; a.err = "round trip mismatch"; break; }
        const double tc = secs(t0, t1), td = secs(t2, t3);
        a.t_c += tc; a.t_d += td; a.chunks++; a.samples += n; a.bytes += comp.size();
        if (n) { a.bits += 8.0 * comp.size() / n; a.mibs_c += 2.0 * n / tc / (1 << 20); a.mibs_d += 2.0 * n / td / (1 << 20); }
    }
    return a;
}

static void print(const char* name, const Acc& a, bool last)
{
    const double k = a.chunks ? (double)a.chunks : 1.0;
    printf("  \"%s\": {\"ok\": %s, \"chunks\": %zu, \"samples\": %zu, \"bits_per_symbol_mean\": %.4f, \"compress_MiBps_mean_of_chunks\": %.1f, "
           "\"decompress_MiBps_mean_of_chunks\": %.1f, \"compress_us_per_chunk\": %.1f, \"decompress_us_per_chunk\": %.1f, "
           "\"compress_GBps_total\": %.3f, \"decompress_GBps_total\": %.3f, \"error\": \"%s\"}%s\n",
           name, a.ok ? "true" : "false", a.chunks, a.samples, a.bits / k, a.mibs_c / k, a.mibs_d / k, 1e6 * a.t_c / k, 1e6 * a.t_d / k,
           a.t_c > 0 ? 2e-9 * a.samples / a.t_c : 0.0, a.t_d > 0 ? 2e-9 * a.samples / a.t_d : 0.0, a.err.c_str(), last ? "" : ",");
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: seam_bench <chunkfile> [threads]\n"); return 2; }
    const std::vector<Chunk> chunks = load(argv[1]);
    const int threads = argc > 2 ? atoi(argv[2]) : 1;
    if (chunks.empty()) { fprintf(stderr, "no chunks in %s\n", argv[1]); return 2; }
    arrow::MemoryPool* pool = arrow::default_memory_pool();
    pod5::ReadData rd;
    pgnano::PGNanoReaderState st;

    // ---- C++ seam (signal_table_writer.cpp:98,107 / signal_table_reader.cpp:133,138) ----
    auto cxx_pg_enc = [&](const std::vector<int16_t>& x, std::vector<uint8_t>& out) -> std::string {
        auto r = pgnano::compress_signal(gsl::make_span(x.data(), x.size()), pool, rd, false);
        if (!r.ok()) return r.status().ToString();
        out.assign((*r)->data(), (*r)->data() + (*r)->size()); return "";
    };
    auto cxx_pg_dec = [&](const std::vector<uint8_t>& in, std::vector<int16_t>& out) -> std::string {
        auto s = pgnano::decompress_signal(gsl::make_span(in.data(), in.size()), pool, gsl::make_span(out.data(), out.size()), st);
        return s.ok() ? "" : s.ToString();
    };
    auto cxx_vbz_enc = [&](const std::vector<int16_t>& x, std::vector<uint8_t>& out) -> std::string {
        auto r = pod5::compress_signal(gsl::make_span(x.data(), x.size()), pool);
        if (!r.ok()) return r.status().ToString();
        out.assign((*r)->data(), (*r)->data() + (*r)->size()); return "";
    };
    auto cxx_vbz_dec = [&](const std::vector<uint8_t>& in, std::vector<int16_t>& out) -> std::string {
        auto s = pod5::decompress_signal(gsl::make_span(in.data(), in.size()), pool, gsl::make_span(out.data(), out.size()));
        return s.ok() ? "" : s.ToString();
    };
    // ---- C API names (benchmark.cpp:145,175) ----
    auto c_enc = [&](bool pg) {
        return [pg](const std::vector<int16_t>& x, std::vector<uint8_t>& out) -> std::string {
            size_t sz = x.size() * 4 + 1024;                              // benchmark.cpp:139 gives the call 4 bytes per sample
            out.resize(sz);
            static const int16_t zero = 0;
            const pod5_error_t e = pg ? pod5_pinanoraw_compress_signal(x.empty() ? &zero : x.data(), x.size(), (char*)out.data(), &sz)
                                      : pod5_vbz_compress_signal(x.empty() ? &zero : x.data(), x.size(), (char*)out.data(), &sz);
            if (e != POD5_OK) return pod5_get_error_string();
            out.resize(sz); return "";
        };
    };
    auto c_dec = [&](bool pg) {
        return [pg](const std::vector<uint8_t>& in, std::vector<int16_t>& out) -> std::string {
            static short dummy[8];
            const pod5_error_t e = pg ? pod5_pinanoraw_decompress_signal((const char*)in.data(), in.size(), out.size(), out.empty() ? dummy : out.data())
                                      : pod5_vbz_decompress_signal((const char*)in.data(), in.size(), out.size(), out.empty() ? dummy : out.data());
            return e == POD5_OK ? "" : pod5_get_error_string();
        };
    };
    // warm-up: context creation, first kernel launches, staging buffers
    { std::vector<Chunk> w(chunks.begin(), chunks.begin() + (chunks.size() < 4 ? chunks.size() : 4)); run(w, cxx_pg_enc, cxx_pg_dec); run(w, cxx_vbz_enc, cxx_vbz_dec); }

    printf("{\n  \"variant\": %d, \"threads\": %d,\n", rnsc_get_pgnano_variant(), threads);
    print("cxx_pgnano", run(chunks, cxx_pg_enc, cxx_pg_dec), false);
    print("cxx_vbz", run(chunks, cxx_vbz_enc, cxx_vbz_dec), false);
    print("c_api_pinanoraw", run(chunks, c_enc(true), c_dec(true)), false);
    print("c_api_vbz", run(chunks, c_enc(false), c_dec(false)), threads <= 1);
    if (threads > 1) {      // concurrent callers, each on its own context (AsyncSignalLoader-style workers): aggregate rate
        std::vector<Acc> res(threads);
        std::vector<std::thread> th;
        auto t0 = Clock::now();
        for (int t = 0; t < threads; ++t) th.emplace_back([&, t] { res[t] = run(chunks, c_enc(true), c_dec(true)); });
        for (auto& x : th) x.join();
        const double wall = secs(t0, Clock::now());
        Acc a; a.ok = true;
        for (auto& r : res) { a.ok = a.ok && r.ok; a.chunks += r.chunks; a.samples += r.samples; a.bytes += r.bytes; a.bits += r.bits; a.mibs_c += r.mibs_c; a.mibs_d += r.mibs_d; if (!r.ok) a.err = r.err; }
        a.t_c = a.t_d = wall / 2;      // total = samples (encoded and decoded once each) over the wall time of both halves together
        print("c_api_pinanoraw_concurrent", a, true);
    }
    printf("}\n");
    return 0;
}
