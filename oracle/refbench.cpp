/*
 * refbench.cpp -- TEST / BENCH INFRASTRUCTURE ONLY.
 *
 * Times a QB3.h implementation loaded with dlopen (normally oracle/_ref/libQB3ref.so, the
 * reference compiled from /root/reference) tile-parallel on the host cores: one encoder or
 * decoder handle per tile, a std::thread pool pulling tile indices from an atomic counter
 * (BASELINE.md 3). The library under test is single-threaded and has no mutable
 * globals, so distinct handles are safe to drive concurrently.
 */
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <dlfcn.h>
#include <thread>
#include <vector>

#define QB3_MAXBANDS 256
#include "QB3.h"

namespace {
struct api {
    void *h = nullptr;
    decltype(&qb3_create_encoder) create_encoder;
    decltype(&qb3_destroy_encoder) destroy_encoder;
    decltype(&qb3_set_encoder_coreband) set_coreband;
    decltype(&qb3_set_encoder_quanta) set_quanta;
    decltype(&qb3_set_encoder_mode) set_mode;
    decltype(&qb3_max_encoded_size) max_size;
    decltype(&qb3_encode) encode;
    decltype(&qb3_read_start) read_start;
    decltype(&qb3_read_info) read_info;
    decltype(&qb3_read_data) read_data;
    decltype(&qb3_destroy_decoder) destroy_decoder;
};

template <typename F> bool sym(void *h, const char *name, F &f)
{
    f = reinterpret_cast<F>(dlsym(h, name));
    return f != nullptr;
}

bool load(const char *path, api &a)
{
    a.h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!a.h) return false;
    return sym(a.h, "qb3_create_encoder", a.create_encoder) && sym(a.h, "qb3_destroy_encoder", a.destroy_encoder)
        && sym(a.h, "qb3_set_encoder_coreband", a.set_coreband) && sym(a.h, "qb3_set_encoder_quanta", a.set_quanta)
        && sym(a.h, "qb3_set_encoder_mode", a.set_mode) && sym(a.h, "qb3_max_encoded_size", a.max_size)
        && sym(a.h, "qb3_encode", a.encode) && sym(a.h, "qb3_read_start", a.read_start)
        && sym(a.h, "qb3_read_info", a.read_info) && sym(a.h, "qb3_read_data", a.read_data)
        && sym(a.h, "qb3_destroy_decoder", a.destroy_decoder);
}

template <typename F> double run_pool(int nthreads, size_t ntiles, F &&body)
{
    std::atomic<size_t> next(0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++)
        pool.emplace_back([&, t]() {
            for (size_t i = next.fetch_add(1); i < ntiles; i = next.fetch_add(1)) body(i, t);
        });
    for (auto &th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
} // namespace

/*
 * src: ntiles tiles back to back, each w*h*bands values of the type. cband may be NULL (library default).
 * dst_streams: ntiles slots of slot_bytes each (>= qb3_max_encoded_size), sizes[ntiles] receives stream lengths.
 * decoded: optional ntiles*tile_bytes buffer; when NULL a per-thread scratch tile is used.
 * Returns 0 on success; enc_s / dec_s are the best of reps wall-clock seconds of the encode / decode pools.
 */
extern "C" __attribute__((visibility("default")))
int refbench_run(const char *libpath, size_t ntiles, size_t w, size_t h, size_t bands, int dtype, int mode,
                 const size_t *cband, uint64_t quanta, const void *src, uint8_t *dst_streams, size_t slot_bytes,
                 size_t *sizes, void *decoded, int nthreads, int reps, double *enc_s, double *dec_s)
{
    api a;
    if (!load(libpath, a)) return 1;
    static const size_t tsz[8] = {1, 1, 2, 2, 4, 4, 8, 8};
    const size_t tile_bytes = w * h * bands * tsz[dtype];
    std::atomic<int> err(0);
    double best_e = 1e30, best_d = 1e30;
    for (int r = 0; r < reps; r++) {
        double te = run_pool(nthreads, ntiles, [&](size_t i, int) {
            encsp e = a.create_encoder(w, h, bands, qb3_dtype(dtype));
            if (!e) { err = 2; return; }
            if (cband) {
                size_t cb[QB3_MAXBANDS];
                memcpy(cb, cband, bands * sizeof(size_t));
                a.set_coreband(e, bands, cb);
            }
            if (quanta > 1) a.set_quanta(e, quanta, false);
            a.set_mode(e, qb3_mode(mode));
            if (a.max_size(e) > slot_bytes) { err = 3; a.destroy_encoder(e); return; }
            sizes[i] = a.encode(e, const_cast<uint8_t *>(static_cast<const uint8_t *>(src)) + i * tile_bytes,
                                dst_streams + i * slot_bytes);
            if (!sizes[i]) err = 4;
            a.destroy_encoder(e);
        });
        if (te < best_e) best_e = te;
    }
    if (err) return err;
    std::vector<std::vector<uint8_t>> scratch(nthreads);
    if (!decoded) for (auto &s : scratch) s.resize(tile_bytes);
    for (int r = 0; r < reps; r++) {
        double td = run_pool(nthreads, ntiles, [&](size_t i, int t) {
            size_t dims[3];
            decsp d = a.read_start(dst_streams + i * slot_bytes, sizes[i], dims);
            if (!d) { err = 5; return; }
            void *out = decoded ? static_cast<uint8_t *>(decoded) + i * tile_bytes : scratch[t].data();
            if (!a.read_info(d) || a.read_data(d, out) != tile_bytes) err = 6;
            a.destroy_decoder(d);
        });
        if (td < best_d) best_d = td;
    }
    *enc_s = best_e;
    *dec_s = best_d;
    return err;
}
