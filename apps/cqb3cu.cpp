/*
 * cqb3cu -- QB3 encode / decode utility on top of the B200 library, the counterpart of the reference's cqb3
 * (cqb3.cpp): the same options with the same meaning (cqb3.cpp:68-88, 100-235), the same mode selection
 * (cqb3.cpp:437-462) and band mix search (cqb3.cpp:561-586). The reference reads PNG / JPEG through libicd, which is
 * not part of its tree; this tool reads and writes binary PNM (P5 / P6, 8 or 16 bit) and headerless raw files instead.
 *
 *   cqb3cu [options] <input> [output]
 *     -v -d -b -f -l -q <n|+n> -r -t -m <b,b,b|x>      as cqb3
 *     -s WxHxB[:u8|i8|u16|i16|u32|i32|u64|i64]          the input (or, with -d, the output) is raw samples
 *     -i                                                print the header of a QB3 file as JSON (wasm/qb3decapi.cpp:60-93)
 *   If the input is a folder, every .ppm / .pgm / .pnm (or, with -d, every .qb3) in it is converted; images of the same
 *   geometry travel through the device as ONE batch (qb3cu_pipe_encode / qb3cu_pipe_decode) instead of one call each.
 *   -m x on the device: the image is uploaded once and encoded with all ten band maps on ten CUDA streams, only the
 *   smallest stream comes back (the reference runs ten full encodes one after the other, cqb3.cpp:561-586).
 */
#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "QB3.h"
#include "qb3cu.h"

using namespace std;
namespace fs = std::filesystem;

struct Options {
    uint64_t quanta = 0;
    string in_fname, out_fname, error, mapping, raw_spec;
    bool best = false, trim = false, rle = false, legacy = false, verbose = false, ftl = false, is_folder = false,
         away = false, decode = false, info = false;
};

struct Image {
    size_t w = 0, h = 0, bands = 0;
    qb3_dtype dt = QB3_U8;
    vector<uint8_t> px; /* native endian samples, band interleaved */
};

static const size_t TSIZE[8] = {1, 1, 2, 2, 4, 4, 8, 8};

static int usage(const Options &opt)
{
    cerr << opt.error << "\n\n"
        "cqb3cu [options] <input_filename> [output_filename]\n"
        "Options:\n"
        "\t-v : verbose\n"
        "\t-d : decode from QB3\n"
        "\t-i : print the QB3 header as JSON\n"
        "\t-s WxHxB[:type] : raw samples instead of PNM (type u8 i8 u16 i16 u32 i32 u64 i64)\n"
        "\n"
        "Compression only options:\n"
        "\t-b : best compression\n"
        "\t-f : fastest compression\n"
        "\t-l : legacy mode (deprecated)\n"
        "\t-q <n> : quanta, +n rounds away from zero\n"
        "\t-r : reverse RLE behavior, off for best, on for fast\n"
        "\t-t : trim input to multiple of 4x4 pixels\n"
        "\t-m <b,b,b> : core band mapping\n"
        "\t-m x : exhaustive band mapping search\n\n"
        "\tIf input is a folder, all .ppm/.pgm/.pnm or (-d) .qb3 files will be processed, same sized ones as one batch\n";
    return 1;
}

static bool isbandmap(const string &s)
{
    return !s.empty() && s.find_first_not_of("0123456789,") == string::npos;
}

static bool parse_args(int argc, char **argv, Options &opt)
{
    for (int i = 1; i < argc; i++) {
        if (argv[i][0] == '-' && argv[i][1] != 0) {
            switch (argv[i][1]) {
            case 'v': opt.verbose = true; break;
            case 'b': opt.best = true; break;
            case 'd': opt.decode = true; break;
            case 'f': opt.ftl = true; break;
            case 't': opt.trim = true; break;
            case 'l': opt.legacy = true; break;
            case 'r': opt.rle = true; break;
            case 'i': opt.info = true; break;
            case 's':
                if (i + 1 >= argc) { opt.error = "-s needs WxHxB[:type]"; return false; }
                opt.raw_spec = argv[++i];
                break;
            case 'q':
                opt.quanta = 2;
                if (i + 1 < argc) {
                    const char c = argv[i + 1][0];
                    if (isdigit((unsigned char)c)) { opt.away = false; opt.quanta = strtoull(argv[++i], nullptr, 10); }
                    else if (c == '+') { opt.away = true; opt.quanta = strtoull(argv[++i] + 1, nullptr, 10); }
                }
                break;
            case 'm':
                opt.mapping = "-";
                if (i + 1 < argc && (string(argv[i + 1]) == "x" || isbandmap(argv[i + 1]))) opt.mapping = argv[++i];
                break;
            default:
                opt.error = "Unknown option provided";
                return false;
            }
        }
        else if (opt.in_fname.empty()) opt.in_fname = argv[i];
        else if (opt.out_fname.empty()) opt.out_fname = argv[i];
        else { opt.error = "Too many positional arguments provided"; return false; }
    }
    if (opt.in_fname.empty()) { opt.error = "Need at least the input file name"; return false; }
    opt.is_folder = fs::is_directory(opt.in_fname);
    if (opt.is_folder && !opt.out_fname.empty() && !fs::is_directory(opt.out_fname)) {
        opt.error = "Output name must be empty or a folder when input is a folder";
        return false;
    }
    if (opt.ftl) opt.best = opt.rle = opt.legacy = false;
    if (opt.decode && (opt.trim || opt.best)) { opt.error = "-t and -b are invalid for QB3 decoding"; return false; }
    return true;
}

static string out_name(const Options &opt, const string &in, const string &ext)
{
    string stem = fs::path(in).stem().string();
    if (opt.is_folder) return (fs::path(opt.out_fname.empty() ? opt.in_fname : opt.out_fname) / (stem + ext)).string();
    if (opt.out_fname.empty()) return stem + ext;
    if (fs::is_directory(opt.out_fname)) return (fs::path(opt.out_fname) / (stem + ext)).string();
    return opt.out_fname;
}

static bool read_file(const string &name, vector<uint8_t> &out)
{
    FILE *f = fopen(name.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    rewind(f);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = n >= 0 && (n == 0 || fread(out.data(), (size_t)n, 1, f) == 1);
    fclose(f);
    return ok;
}

static bool write_file(const string &name, const void *p, size_t n)
{
    FILE *f = fopen(name.c_str(), "wb");
    if (!f) return false;
    const bool ok = n == 0 || fwrite(p, n, 1, f) == 1;
    fclose(f);
    return ok;
}

/* "WxHxB[:type]" */
static bool parse_raw_spec(const string &s, Image &im)
{
    static const char *names[8] = {"u8", "i8", "u16", "i16", "u32", "i32", "u64", "i64"};
    unsigned long w = 0, h = 0, b = 0;
    char type[8] = "u8";
    const int n = sscanf(s.c_str(), "%lux%lux%lu:%7s", &w, &h, &b, type);
    if (n < 3) return false;
    im.w = w; im.h = h; im.bands = b;
    for (int i = 0; i < 8; i++)
        if (!strcmp(type, names[i])) { im.dt = (qb3_dtype)i; return true; }
    return false;
}

/* binary PNM: P5 (one band) or P6 (three), maxval up to 65535, 16 bit samples big endian in the file */
static bool parse_pnm(const vector<uint8_t> &f, Image &im)
{
    if (f.size() < 8 || f[0] != 'P' || (f[1] != '5' && f[1] != '6')) return false;
    size_t pos = 2;
    unsigned long v[3];
    for (int k = 0; k < 3; k++) {
        for (;;) { /* white space and comments */
            while (pos < f.size() && isspace(f[pos])) pos++;
            if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') pos++; continue; }
            break;
        }
        if (pos >= f.size() || !isdigit(f[pos])) return false;
        v[k] = 0;
        while (pos < f.size() && isdigit(f[pos])) v[k] = v[k] * 10 + (f[pos++] - '0');
    }
    pos++; /* the single white space after maxval */
    im.w = v[0]; im.h = v[1]; im.bands = f[1] == '5' ? 1 : 3;
    if (v[2] < 1 || v[2] > 65535) return false;
    im.dt = v[2] > 255 ? QB3_U16 : QB3_U8;
    const size_t n = im.w * im.h * im.bands, ts = TSIZE[im.dt];
    if (f.size() < pos + n * ts) return false;
    im.px.resize(n * ts);
    if (ts == 1) memcpy(im.px.data(), f.data() + pos, n);
    else {
        uint16_t *d = reinterpret_cast<uint16_t *>(im.px.data());
        for (size_t i = 0; i < n; i++) d[i] = (uint16_t)(f[pos + 2 * i] << 8 | f[pos + 2 * i + 1]);
    }
    return true;
}

static bool write_image(const string &name, const Image &im, bool raw)
{
    const size_t n = im.w * im.h * im.bands;
    if (raw || !(im.dt == QB3_U8 || im.dt == QB3_U16) || !(im.bands == 1 || im.bands == 3))
        return write_file(name, im.px.data(), im.px.size());
    char hdr[64];
    const int hl = snprintf(hdr, sizeof(hdr), "P%c\n%zu %zu\n%d\n", im.bands == 1 ? '5' : '6', im.w, im.h, im.dt == QB3_U8 ? 255 : 65535);
    vector<uint8_t> out(hl + im.px.size());
    memcpy(out.data(), hdr, hl);
    if (im.dt == QB3_U8) memcpy(out.data() + hl, im.px.data(), n);
    else {
        const uint16_t *s = reinterpret_cast<const uint16_t *>(im.px.data());
        for (size_t i = 0; i < n; i++) { out[hl + 2 * i] = (uint8_t)(s[i] >> 8); out[hl + 2 * i + 1] = (uint8_t)s[i]; }
    }
    return write_file(name, out.data(), out.size());
}

static bool load_image(const Options &opt, const string &name, Image &im)
{
    vector<uint8_t> f;
    if (!read_file(name, f)) { cerr << "Can't read " << name << "\n"; return false; }
    if (!opt.raw_spec.empty()) {
        if (!parse_raw_spec(opt.raw_spec, im)) { cerr << "Bad -s specification\n"; return false; }
        if (f.size() < im.w * im.h * im.bands * TSIZE[im.dt]) { cerr << name << " is shorter than -s says\n"; return false; }
        f.resize(im.w * im.h * im.bands * TSIZE[im.dt]);
        im.px.swap(f);
        return true;
    }
    if (!parse_pnm(f, im)) { cerr << name << " is not a binary PNM (P5 / P6)\n"; return false; }
    return true;
}

/* the mode the options ask for (cqb3.cpp:437-462) */
static qb3_mode pick_mode(const Options &opt)
{
    qb3_mode mode = opt.best ? QB3M_BEST : QB3M_BASE;
    if (opt.legacy) mode = mode == QB3M_BEST ? QB3M_CF_RLE : QB3M_BASE_Z;
    if (opt.rle) {
        if (mode == QB3M_BEST) mode = QB3M_CF_H;
        else if (mode == QB3M_BASE) mode = QB3M_RLE_H;
        else if (mode == QB3M_BASE_Z) mode = QB3M_RLE;
        else if (mode == QB3M_CF_RLE) mode = QB3M_CF;
    }
    if (opt.ftl) mode = QB3M_FTL;
    return mode;
}

/* -t: drop the first column / line when the remainder modulo four is above one, then cut to multiples of four
   (cqb3.cpp:394-404); returns the byte offset of the first pixel kept, the stride stays the original line */
static size_t trim(const Options &opt, Image &im)
{
    size_t offset = 0;
    if (opt.trim && (im.w % 4 || im.h % 4)) {
        const size_t ts = TSIZE[im.dt], stride = im.w * im.bands;
        if (im.w % 4 > 1) offset += ts * im.bands;
        if (im.h % 4 > 1) offset += stride * ts;
        im.w -= im.w % 4;
        im.h -= im.h % 4;
        cout << "Trimmed to " << im.w << "x" << im.h << endl;
    }
    return offset;
}

/* band map from "b,b,b" the way cqb3.cpp:413-429 reads it; missing entries are identity */
static void parse_mapping(const string &m, size_t bands, size_t *bmap)
{
    string rest = m == "-" ? "" : m;
    for (size_t i = 0; i < bands; i++) {
        if (rest.empty()) { bmap[i] = i; continue; }
        char *end = nullptr;
        bmap[i] = strtoul(rest.c_str(), &end, 10);
        while (*end == ',') end++;
        rest = end;
    }
}

/* One image through the QB3.h API, exactly the calls cqb3 makes (cqb3.cpp:405-481). */
static int encode_one(const Options &opt, Image im, const string &mapping, vector<uint8_t> &dest, double &seconds)
{
    const size_t stride = im.w * im.bands; /* of the untrimmed image */
    const size_t offset = trim(opt, im);
    encsp q = qb3_create_encoder(im.w, im.h, im.bands, im.dt);
    if (!q) { cerr << "Can't create the encoder (geometry, type, or no CUDA device)\n"; return 1; }
    qb3_set_encoder_stride(q, stride);
    dest.resize(qb3_max_encoded_size(q));
    if (!mapping.empty()) {
        size_t bmap[QB3_MAXBANDS];
        parse_mapping(mapping, im.bands, bmap);
        if (!qb3_set_encoder_coreband(q, im.bands, bmap)) cerr << "Invalid band mapping, adjusted\n";
    }
    const qb3_mode mode = pick_mode(opt);
    int rc = 0;
    if (mode != qb3_set_encoder_mode(q, mode)) { cerr << "Invalid mode\n"; rc = 1; }
    if (!rc && opt.quanta > 1 && !qb3_set_encoder_quanta(q, opt.quanta, opt.away)) { cerr << "Invalid quanta\n"; rc = 1; }
    if (!rc) {
        const auto t1 = chrono::high_resolution_clock::now();
        const size_t n = qb3_encode(q, im.px.data() + offset, dest.data());
        seconds += chrono::duration<double>(chrono::high_resolution_clock::now() - t1).count();
        if (n == 0) { cerr << "Encoding failed, state " << qb3_get_encoder_state(q) << "\n"; rc = 2; }
        dest.resize(n);
    }
    qb3_destroy_encoder(q);
    return rc;
}

static void fill_config(const Options &opt, const Image &im, const string &mapping, qb3cu_config &cfg)
{
    qb3cu_config_init(&cfg, (uint32_t)im.w, (uint32_t)im.h, (uint32_t)im.bands, (uint32_t)im.dt);
    cfg.mode = pick_mode(opt);
    cfg.quanta = opt.quanta > 1 ? opt.quanta : 1;
    cfg.away = opt.away;
    if (!mapping.empty()) {
        /* what qb3_set_encoder_coreband makes of the list (QB3encode.cpp:63-77): out of range means self, a band
           that is referred to becomes a core band, in band order */
        size_t bmap[QB3_MAXBANDS];
        parse_mapping(mapping, im.bands, bmap);
        for (size_t i = 0; i < im.bands; i++) cfg.cband[i] = (uint8_t)(bmap[i] < im.bands ? bmap[i] : i);
        for (size_t i = 0; i < im.bands; i++)
            if (cfg.cband[i] != i) cfg.cband[cfg.cband[i]] = cfg.cband[i];
    }
}

/* -m x: the ten RGB band maps of cqb3.cpp:567-570 on the device. The image goes up once; every map is only measured
   (qb3cu_encoded_size_batch: the encode kernel without its packing, no destination), each on a stream of its own;
   the first of the smallest is then encoded for real and only that stream comes back. */
static int encode_bandmix(const Options &opt, Image im, vector<uint8_t> &dest, double &seconds)
{
    static const char *combos[10] = {"1,1,1", "0,0,0", "0,0,2", "0,1,0", "0,1,1", "0,1,2", "0,2,2", "1,1,2", "2,1,2", "2,2,2"};
    const size_t stride = im.w * im.bands, offset = trim(opt, im), ts = TSIZE[im.dt];
    qb3cu_config cfg[10];
    for (int k = 0; k < 10; k++) { fill_config(opt, im, combos[k], cfg[k]); cfg[k].stride = stride; }
    const size_t slot = qb3cu_slot_bytes(&cfg[0]), extent = ((im.h - 1) * stride + im.w * im.bands) * ts;
    uint8_t *d_src = nullptr, *d_dst = nullptr;
    uint64_t *d_sizes = nullptr, sizes[11];
    cudaStream_t st[10];
    if (cudaMalloc(&d_src, extent) || cudaMalloc(&d_dst, slot) || cudaMalloc(&d_sizes, 88)) { cerr << "No device memory\n"; return 2; }
    const auto t1 = chrono::high_resolution_clock::now();
    cudaMemcpy(d_src, im.px.data() + offset, extent, cudaMemcpyHostToDevice);
    int rc = 0;
    for (int k = 0; k < 10; k++) {
        cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking);
        rc |= qb3cu_encoded_size_batch(&cfg[k], d_src, extent, d_sizes + k, 1, st[k]);
    }
    for (int k = 0; k < 10; k++) { cudaStreamSynchronize(st[k]); cudaStreamDestroy(st[k]); }
    cudaMemcpy(sizes, d_sizes, 80, cudaMemcpyDeviceToHost);
    int best = 0;
    for (int k = 0; k < 10 && !rc; k++) { /* the first of the smallest, as the reference's strict comparison keeps it */
        if (opt.verbose && (k == 0 || sizes[k] < sizes[best])) cout << "Band mix " << combos[k] << ", size " << sizes[k] << endl;
        if (sizes[k] < sizes[best]) best = k;
    }
    if (!rc) rc = qb3cu_encode_batch(&cfg[best], d_src, extent, d_dst, slot, d_sizes + 10, nullptr, nullptr, 1, nullptr);
    if (!rc) {
        cudaMemcpy(sizes + 10, d_sizes + 10, 8, cudaMemcpyDeviceToHost);
        if (sizes[10] != sizes[best]) { cerr << "Size pass and encode disagree\n"; rc = 2; }
        dest.resize(sizes[10]);
        cudaMemcpy(dest.data(), d_dst, sizes[10], cudaMemcpyDeviceToHost);
    }
    seconds += chrono::duration<double>(chrono::high_resolution_clock::now() - t1).count();
    cudaFree(d_src); cudaFree(d_dst); cudaFree(d_sizes);
    if (rc) cerr << "Encoding failed\n";
    return rc ? 2 : 0;
}

static int encode_file(const Options &opt, const string &in, const string &out)
{
    Image im;
    if (!load_image(opt, in, im)) return 1;
    if (opt.verbose) cout << "Input " << im.w << "x" << im.h << "@" << im.bands << (TSIZE[im.dt] > 1 ? " 16bit or more\n" : "\n");
    vector<uint8_t> dest;
    double seconds = 0;
    const size_t raw = im.px.size();
    int rc;
    if (opt.mapping == "x" && (im.bands == 3 || im.bands == 4)) rc = encode_bandmix(opt, im, dest, seconds);
    else if (opt.mapping == "x" && im.bands > 4) { cerr << "Exhaustive band mix implemented only for RGB/RGBA inputs\n"; return 1; }
    else rc = encode_one(opt, im, opt.mapping == "x" ? "" : opt.mapping, dest, seconds);
    if (rc) return rc;
    if (opt.verbose)
        cout << "Output\nSize: " << dest.size() << "\nEncode time : " << seconds << "s\nRatio " << dest.size() * 100.0 / raw
             << "%, rate : " << raw / seconds / 1024 / 1024 << " MB/s\n";
    if (!write_file(out, dest.data(), dest.size())) { cerr << "Can't write " << out << "\n"; return 1; }
    return 0;
}

static int decode_file(const Options &opt, const string &in, const string &out)
{
    vector<uint8_t> src;
    if (!read_file(in, src)) { cerr << "Can't read " << in << "\n"; return 1; }
    size_t sz[3];
    decsp q = qb3_read_start(src.data(), src.size(), sz);
    if (!q) { cerr << in << " is not a QB3 stream\n"; return 1; }
    if (!qb3_read_info(q)) { cerr << "Can't read the QB3 headers\n"; qb3_destroy_decoder(q); return 1; }
    Image im;
    im.w = sz[0]; im.h = sz[1]; im.bands = sz[2]; im.dt = qb3_get_type(q);
    im.px.resize(qb3_decoded_size(q));
    const auto t1 = chrono::high_resolution_clock::now();
    const size_t n = qb3_read_data(q, im.px.data());
    const double seconds = chrono::duration<double>(chrono::high_resolution_clock::now() - t1).count();
    qb3_destroy_decoder(q);
    if (!n) { cerr << "Decoding failed\n"; return 2; }
    if (opt.verbose)
        cout << "Image " << im.w << "x" << im.h << "@" << im.bands << "\nDecode time : " << seconds << "s, rate : "
             << im.px.size() / seconds / 1024 / 1024 << " MB/s\n";
    if (!write_image(out, im, !opt.raw_spec.empty())) { cerr << "Can't write " << out << "\n"; return 1; }
    return 0;
}

/* -i: the header as JSON, the fields and names of wasm/qb3decapi.cpp:60-93. Needs no device. */
static int info_file(const string &in)
{
    static const char *types[8] = {"uint8", "int8", "uint16", "int16", "uint32", "int32", "uint64", "int64"};
    vector<uint8_t> src;
    if (!read_file(in, src)) { cerr << "Can't read " << in << "\n"; return 1; }
    size_t sz[3] = {0, 0, 0};
    decsp q = qb3_read_start(src.data(), src.size(), sz);
    if (!q) { cout << "{\"error\": \"Invalid QB3 format\"}\n"; return 1; }
    if (!qb3_read_info(q)) { cout << "{\"error\": \"Failed to read QB3 info\"}\n"; qb3_destroy_decoder(q); return 1; }
    const qb3_mode m = qb3_get_mode(q);
    const char *mode = m == QB3M_BASE_Z ? "base_z" : m == QB3M_CF ? "cf" : m == QB3M_RLE ? "rle" : m == QB3M_CF_RLE ? "cf_rle"
                     : m == QB3M_BASE_H ? "base" : m == QB3M_CF_H ? "cf_h" : m == QB3M_RLE_H ? "rle_h" : m == QB3M_CF_RLE_H ? "best"
                     : m == QB3M_FTL ? "ftl" : m == QB3M_STORED ? "stored" : "invalid";
    cout << "{\"xsize\": " << sz[0] << ", \"ysize\": " << sz[1] << ", \"nbands\": " << sz[2] << ", \"dtype\": \""
         << types[qb3_get_type(q) & 7] << "\", \"mode\": \"" << mode << "\"";
    if (qb3_get_quanta(q) > 1) cout << ", \"quanta\": " << qb3_get_quanta(q);
    size_t cband[QB3_MAXBANDS] = {0};
    if (qb3_get_coreband(q, cband)) {
        cout << ", \"bandmap\": [";
        for (size_t i = 0; i < sz[2]; i++) cout << (i ? ", " : "") << cband[i];
        cout << "]";
    }
    else cout << ", \"bandmap\": null";
    cout << "}\n";
    qb3_destroy_decoder(q);
    return 0;
}

/* folder mode: same sized images as one batch through the host pipeline */
static int encode_folder(const Options &opt)
{
    struct Key { size_t w, h, b; int dt; bool operator<(const Key &o) const { return tie(w, h, b, dt) < tie(o.w, o.h, o.b, o.dt); } };
    map<Key, vector<pair<string, Image>>> groups;
    for (auto &e : fs::directory_iterator(opt.in_fname)) {
        string ext = e.path().extension().string();
        transform(ext.begin(), ext.end(), ext.begin(), ::tolower);
        if (ext != ".ppm" && ext != ".pgm" && ext != ".pnm") continue;
        Image im;
        if (!load_image(opt, e.path().string(), im)) continue;
        groups[Key{im.w, im.h, im.bands, (int)im.dt}].emplace_back(e.path().string(), std::move(im));
    }
    int rc = 0;
    for (auto &g : groups) {
        auto &items = g.second;
        const size_t n = items.size(), tile = items[0].second.px.size();
        qb3cu_config cfg;
        fill_config(opt, items[0].second, opt.mapping == "x" ? "" : opt.mapping, cfg);
        const size_t slot = qb3cu_slot_bytes(&cfg);
        uint8_t *h_src = static_cast<uint8_t *>(qb3cu_host_alloc(n * tile)), *h_packed = static_cast<uint8_t *>(qb3cu_host_alloc(n * slot));
        qb3cu_pipe *pipe = qb3cu_pipe_create(&cfg, 0, 0);
        vector<uint64_t> offs(n), sizes(n);
        uint64_t total = 0;
        if (!h_src || !h_packed || !pipe) { cerr << "Can't set the batch up (no CUDA device?)\n"; return 2; }
        for (size_t i = 0; i < n; i++) memcpy(h_src + i * tile, items[i].second.px.data(), tile);
        const auto t1 = chrono::high_resolution_clock::now();
        const int r = qb3cu_pipe_encode(pipe, h_src, tile, h_packed, n * slot, offs.data(), sizes.data(), &total, n);
        const double s = chrono::duration<double>(chrono::high_resolution_clock::now() - t1).count();
        if (r) { cerr << "Batch encode failed\n"; rc = 2; }
        for (size_t i = 0; i < n && !r; i++)
            if (!write_file(out_name(opt, items[i].first, ".qb3"), h_packed + offs[i], sizes[i])) rc = 1;
        if (opt.verbose)
            cout << n << " images " << g.first.w << "x" << g.first.h << "@" << g.first.b << ": " << total << " bytes, "
                 << n * tile / s / 1024 / 1024 << " MB/s\n";
        qb3cu_pipe_destroy(pipe);
        qb3cu_host_free(h_src); qb3cu_host_free(h_packed);
    }
    return rc;
}

/* folder mode, decoding: streams of the same geometry and type as one batch through the host pipeline */
static int decode_folder(const Options &opt)
{
    struct Item { string name; vector<uint8_t> bytes; };
    struct Key { size_t w, h, b; int dt; bool operator<(const Key &o) const { return tie(w, h, b, dt) < tie(o.w, o.h, o.b, o.dt); } };
    map<Key, vector<Item>> groups;
    int rc = 0;
    for (auto &e : fs::directory_iterator(opt.in_fname)) {
        string ext = e.path().extension().string();
        transform(ext.begin(), ext.end(), ext.begin(), ::tolower);
        if (ext != ".qb3") continue;
        Item it;
        it.name = e.path().string();
        size_t sz[3];
        decsp q = read_file(it.name, it.bytes) ? qb3_read_start(it.bytes.data(), it.bytes.size(), sz) : nullptr;
        if (!q || !qb3_read_info(q)) { cerr << it.name << " is not a QB3 stream\n"; rc = 1; if (q) qb3_destroy_decoder(q); continue; }
        const Key k{sz[0], sz[1], sz[2], (int)qb3_get_type(q)};
        qb3_destroy_decoder(q);
        groups[k].push_back(std::move(it));
    }
    const string ext_out = opt.raw_spec.empty() ? ".pnm" : ".raw";
    for (auto &g : groups) {
        auto &items = g.second;
        const size_t n = items.size();
        Image im;
        im.w = g.first.w; im.h = g.first.h; im.bands = g.first.b; im.dt = (qb3_dtype)g.first.dt;
        const size_t tile = im.w * im.h * im.bands * TSIZE[im.dt];
        qb3cu_config cfg;
        qb3cu_config_init(&cfg, (uint32_t)im.w, (uint32_t)im.h, (uint32_t)im.bands, (uint32_t)im.dt);
        cfg.mode = QB3M_BEST; /* only a hint for the decoder: RLE streams may be among them */
        vector<uint64_t> offs(n), lens(n);
        uint64_t total = 0;
        for (size_t i = 0; i < n; i++) { offs[i] = total; lens[i] = items[i].bytes.size(); total += (lens[i] + 15) & ~(uint64_t)15; }
        uint8_t *h_streams = static_cast<uint8_t *>(qb3cu_host_alloc(total + 16)), *h_px = static_cast<uint8_t *>(qb3cu_host_alloc(n * tile));
        qb3cu_pipe *pipe = qb3cu_pipe_create(&cfg, 0, 0);
        if (!h_streams || !h_px || !pipe) { cerr << "Can't set the batch up (no CUDA device?)\n"; return 2; }
        for (size_t i = 0; i < n; i++) memcpy(h_streams + offs[i], items[i].bytes.data(), lens[i]);
        vector<uint32_t> status(n, 0);
        const auto t1 = chrono::high_resolution_clock::now();
        const int r = qb3cu_pipe_decode(pipe, h_streams, offs.data(), lens.data(), h_px, tile, status.data(), 0, n);
        const double s = chrono::duration<double>(chrono::high_resolution_clock::now() - t1).count();
        if (r) { cerr << "Batch decode failed\n"; rc = 2; }
        for (size_t i = 0; i < n && !r; i++) {
            if (status[i] != QB3CU_TILE_OK) { cerr << items[i].name << ": decoding failed\n"; rc = 2; continue; }
            im.px.assign(h_px + i * tile, h_px + (i + 1) * tile);
            if (!write_image(out_name(opt, items[i].name, ext_out), im, !opt.raw_spec.empty())) rc = 1;
        }
        if (opt.verbose)
            cout << n << " streams " << im.w << "x" << im.h << "@" << im.bands << ": " << n * tile / s / 1024 / 1024 << " MB/s\n";
        qb3cu_pipe_destroy(pipe);
        qb3cu_host_free(h_streams); qb3cu_host_free(h_px);
    }
    return rc;
}

int main(int argc, char **argv)
{
    Options opt;
    if (!parse_args(argc, argv, opt)) return usage(opt);
    qb3cu_api_max_bands(QB3_MAXBANDS); /* this tool is compiled with -DQB3_MAXBANDS=256: its band maps hold that many */
    if (opt.info) return info_file(opt.in_fname);
    if (opt.is_folder) return opt.decode ? decode_folder(opt) : encode_folder(opt);
    if (opt.decode) return decode_file(opt, opt.in_fname, out_name(opt, opt.in_fname, opt.raw_spec.empty() ? ".pnm" : ".raw"));
    return encode_file(opt, opt.in_fname, out_name(opt, opt.in_fname, ".qb3"));
}
