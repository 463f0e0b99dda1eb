/*
 * refflac.c -- hosts the reference's OWN decoder binary on Linux so it can serve as the primary oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/flac_oracle.h).  The reference ships its codec only as
 *   /root/reference/Library/BirdNest.Audio/LibFLACDLL/LibFlac.dll   (libFLAC 1.2.1, PE32 i386)
 * and drives it with exactly this call sequence (Library/BirdNest.Audio/FLACDecoder.cs:49-70,207-224,294-300):
 *   FLAC__stream_decoder_new -> init_stream(read, seek, tell, length, eof, write, metadata, error)
 *   -> process_until_end_of_metadata -> { get_state < EndOfStream ? process_single }* -> finish -> delete
 * This freestanding 32-bit program (no libc; built with gcc -m32 -nostdlib, see Makefile) maps the DLL at its
 * ImageBase, binds its 33 MSVCRT + 2 KERNEL32 imports to the small shims below, and calls its exports.
 * The write callback restates the C# interleave (FLACDecoder.cs:552-576 / FLACFileReader.cs:214-243).
 * Nothing of the reference is copied: the DLL is loaded from where it lies (or from oracle/_ref/, a git-ignored
 * binary copy made by the Makefile so the oracle can travel to the GPU box).
 *
 *   refflac dec  <dll> in.flac out.pcm        decode, write interleaved LE PCM, print frames/errors/state
 *   refflac decogg <dll> in.oga out.pcm       the same through FLAC__stream_decoder_init_ogg_stream (Ogg FLAC pages)
 *   refflac bench <dll> in.flac iters         time `iters` full decodes (process_single loop + interleave), print ms each
 *   refflac enc  <dll> ch bps sr bs maxlpc minpo maxpo midside exhaustive in.pcm out.flac   encode packed LE PCM
 */
typedef unsigned char u8; typedef unsigned short u16; typedef unsigned int u32; typedef int i32;
typedef unsigned long long u64; typedef long long i64; typedef u32 size_t;
#define ALIGNED __attribute__((force_align_arg_pointer))
#define NULL ((void*)0)

/* ------------------------------------------------------------------ syscalls (int 0x80, i386) */
static i32 sc3(i32 n, i32 a, i32 b, i32 c) { i32 r; __asm__ volatile("int $0x80" : "=a"(r) : "a"(n), "b"(a), "c"(b), "d"(c) : "memory"); return r; }
static void sys_exit(int c) { sc3(1, c, 0, 0); for (;;) {} }
static i32 sys_read(int fd, void* p, u32 n) { return sc3(3, fd, (i32)p, (i32)n); }
static i32 sys_write(int fd, const void* p, u32 n) { return sc3(4, fd, (i32)p, (i32)n); }
static i32 sys_open(const char* path, int flags, int mode) { return sc3(5, (i32)path, flags, mode); }
static i32 sys_close(int fd) { return sc3(6, fd, 0, 0); }
struct mm { u32 addr, len, prot, flags, fd, off; };
static void* xmmap(u32 addr, u32 len, u32 prot, u32 flags) { struct mm m = {addr, len, prot, flags, (u32)-1, 0}; return (void*)sc3(90, (i32)&m, 0, 0); }
struct ts { i32 sec, nsec; };
static u64 udiv64(u64 n, u32 d) { u64 q = 0, r = 0; for (int i = 63; i >= 0; i--) { r = (r << 1) | ((n >> i) & 1); if (r >= d) { r -= d; q |= 1ull << i; } } return q; }
static u64 now_ns(void) { struct ts t; sc3(265, 1, (i32)&t, 0); return (u64)(u32)t.sec * 1000000000ull + (u32)t.nsec; }

/* ------------------------------------------------------------------ tiny libc */
void* memcpy(void* d, const void* s, size_t n) { u8* a = d; const u8* b = s; while (n--) *a++ = *b++; return d; }
void* memset(void* d, int c, size_t n) { u8* a = d; while (n--) *a++ = (u8)c; return d; }
void* memmove(void* d, const void* s, size_t n) { u8* a = d; const u8* b = s; if (a < b) while (n--) *a++ = *b++; else { a += n; b += n; while (n--) *--a = *--b; } return d; }
static size_t xstrlen(const char* s) { size_t n = 0; while (s[n]) n++; return n; }
static int xstrcmp(const char* a, const char* b) { while (*a && *a == *b) { a++; b++; } return (u8)*a - (u8)*b; }
static void put(const char* s) { sys_write(1, s, xstrlen(s)); }
static void putu(u32 v) { char b[12]; int i = 11; b[i] = 0; do { b[--i] = (char)('0' + v % 10); v /= 10; } while (v); put(b + i); }
static void puthex8(u8 v) { const char* h = "0123456789abcdef"; char b[3] = {h[v >> 4], h[v & 15], 0}; put(b); }
static u32 atou(const char* s) { u32 v = 0; while (*s >= '0' && *s <= '9') v = v * 10 + (u32)(*s++ - '0'); return v; }
static void die(const char* m) { put("refflac: "); put(m); put("\n"); sys_exit(3); }

/* ------------------------------------------------------------------ heap for the DLL (bump allocator, size header) */
static u8* heap; static u32 heap_off, heap_cap;
ALIGNED static void* s_malloc(size_t n) {
    if (!n) n = 1;
    n = (n + 15) & ~15u;
    if (heap_off + n + 16 > heap_cap) die("heap exhausted");
    u32* h = (u32*)(heap + heap_off); h[0] = n; heap_off += n + 16;
    return (u8*)h + 16;
}
ALIGNED static void* s_calloc(size_t a, size_t b) { size_t n = a * b; void* p = s_malloc(n); memset(p, 0, n); return p; }
ALIGNED static void s_free(void* p) { (void)p; }
ALIGNED static void* s_realloc(void* p, size_t n) {
    if (!p) return s_malloc(n);
    u32 old = ((u32*)((u8*)p - 16))[0];
    if (n <= old) return p;
    void* q = s_malloc(n); memcpy(q, p, old); return q;
}
ALIGNED static void* s_memmove(void* d, const void* s, size_t n) { return memmove(d, s, n); }
ALIGNED static void* s_memchr(const void* s, int c, size_t n) { const u8* p = s; while (n--) { if (*p == (u8)c) return (void*)p; p++; } return NULL; }
ALIGNED static char* s_strchr(const char* s, int c) { for (;; s++) { if (*s == (char)c) return (char*)s; if (!*s) return NULL; } }
ALIGNED static char* s_strrchr(const char* s, int c) { const char* r = NULL; for (;; s++) { if (*s == (char)c) r = s; if (!*s) return (char*)r; } }
ALIGNED static int s_strncmp(const char* a, const char* b, size_t n) { while (n--) { if (*a != *b || !*a) return (u8)*a - (u8)*b; a++; b++; } return 0; }
static int lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }
ALIGNED static int s_strnicmp(const char* a, const char* b, size_t n) { while (n--) { int x = lower((u8)*a), y = lower((u8)*b); if (x != y || !x) return x - y; a++; b++; } return 0; }
ALIGNED static char* s_strdup(const char* s) { size_t n = xstrlen(s) + 1; char* p = s_malloc(n); memcpy(p, s, n); return p; }
static int g_errno;
ALIGNED static int* s_errno(void) { return &g_errno; }
ALIGNED static void s_initterm(void (**a)(void), void (**b)(void)) { for (; a < b; a++) if (*a) (*a)(); }
ALIGNED static double s_frexp(double x, int* e) {
    union { double d; u64 u; } v; v.d = x; int ex = (int)((v.u >> 52) & 0x7ff);
    if (ex == 0) { if (x == 0) { *e = 0; return x; } v.d = x * 18014398509481984.0; ex = (int)((v.u >> 52) & 0x7ff); *e = ex - 1022 - 54; }
    else if (ex == 0x7ff) { *e = 0; return x; } else *e = ex - 1022;
    v.u = (v.u & ~(0x7ffULL << 52)) | (1022ULL << 52); return v.d;
}
ALIGNED static double s_strtod(const char* s, char** end) {
    double v = 0, sc = 1; int neg = 0;
    while (*s == ' ') s++;
    if (*s == '-') { neg = 1; s++; } else if (*s == '+') s++;
    while (*s >= '0' && *s <= '9') v = v * 10 + (*s++ - '0');
    if (*s == '.') { s++; while (*s >= '0' && *s <= '9') { sc /= 10; v += (*s++ - '0') * sc; } }
    if (end) *end = (char*)s;
    return neg ? -v : v;
}
ALIGNED static void s_qsort(void* base, size_t n, size_t sz, int (*cmp)(const void*, const void*)) {
    u8* b = base; u8 tmp[256];
    if (sz > sizeof tmp) die("qsort element too large");
    for (size_t i = 1; i < n; i++) {
        memcpy(tmp, b + i * sz, sz);
        size_t j = i;
        while (j > 0 && cmp(b + (j - 1) * sz, tmp) > 0) { memcpy(b + j * sz, b + (j - 1) * sz, sz); j--; }
        memcpy(b + j * sz, tmp, sz);
    }
}
/* MSVC _ftol: ST(0) -> edx:eax, truncating */
void s_ftol(void);
__asm__(".text\n.globl s_ftol\ns_ftol:\n push %ebp\n mov %esp,%ebp\n sub $16,%esp\n fnstcw -2(%ebp)\n movw -2(%ebp),%ax\n orb $0x0c,%ah\n movw %ax,-4(%ebp)\n"
        " fldcw -4(%ebp)\n fistpll -16(%ebp)\n fldcw -2(%ebp)\n mov -16(%ebp),%eax\n mov -12(%ebp),%edx\n leave\n ret\n");
__attribute__((stdcall)) ALIGNED static int k_DisableThreadLibraryCalls(void* h) { (void)h; return 1; }
__attribute__((stdcall)) ALIGNED static void* k_SetUnhandledExceptionFilter(void* f) { (void)f; return NULL; }
static int g_adjust_fdiv = 0; static u8 g_iob[96];
static const char* trap_name = "?";
ALIGNED static void s_trap(void) { put("refflac: DLL called an unbound import (file API?)\n"); sys_exit(4); }

struct shim { const char* name; void* fn; };
static const struct shim shims[] = {
    {"malloc", s_malloc}, {"calloc", s_calloc}, {"free", s_free}, {"realloc", s_realloc}, {"memmove", s_memmove}, {"memchr", s_memchr},
    {"strchr", s_strchr}, {"strrchr", s_strrchr}, {"strncmp", s_strncmp}, {"_strnicmp", s_strnicmp}, {"_strdup", s_strdup}, {"_errno", s_errno},
    {"_initterm", s_initterm}, {"frexp", s_frexp}, {"strtod", s_strtod}, {"qsort", s_qsort}, {"_ftol", s_ftol},
    {"DisableThreadLibraryCalls", k_DisableThreadLibraryCalls}, {"SetUnhandledExceptionFilter", k_SetUnhandledExceptionFilter},
    {"_adjust_fdiv", &g_adjust_fdiv}, {"_iob", g_iob}, {NULL, NULL}};
static void* lookup(const char* nm) { for (const struct shim* s = shims; s->name; s++) if (!xstrcmp(s->name, nm)) return s->fn; (void)trap_name; return (void*)s_trap; }

/* ------------------------------------------------------------------ PE loader */
static u32 rd32(const u8* p) { return p[0] | p[1] << 8 | p[2] << 16 | (u32)p[3] << 24; }
static u16 rd16(const u8* p) { return (u16)(p[0] | p[1] << 8); }
static u8* img; static u32 exp_rva;
static u8 dllfile[262144];
static void pe_load(const char* path) {
    int fd = sys_open(path, 0, 0); if (fd < 0) die("cannot open LibFlac.dll");
    u32 n = 0; for (;;) { i32 r = sys_read(fd, dllfile + n, sizeof dllfile - n); if (r <= 0) break; n += (u32)r; }
    sys_close(fd);
    if (n < 1024 || dllfile[0] != 'M' || dllfile[1] != 'Z') die("not a PE file");
    u8* nt = dllfile + rd32(dllfile + 0x3c), *opt = nt + 24;
    u32 base = rd32(opt + 28), size = rd32(opt + 56), hdrs = rd32(opt + 60); u16 nsec = rd16(nt + 6), optsz = rd16(nt + 20);
    img = xmmap(base, (size + 4095) & ~4095u, 7, 0x2 | 0x10 | 0x20);
    if ((u32)img != base) die("cannot map the DLL at its ImageBase");
    memcpy(img, dllfile, hdrs);
    u8* sec = opt + optsz;
    for (int i = 0; i < nsec; i++, sec += 40) { u32 va = rd32(sec + 12), rawsz = rd32(sec + 16), raw = rd32(sec + 20); if (rawsz) memcpy(img + va, dllfile + raw, rawsz); }
    exp_rva = rd32(opt + 96);
    u32 imp_rva = rd32(opt + 104);
    for (u8* d = img + imp_rva; rd32(d + 12); d += 20) {
        u32 oft = rd32(d) ? rd32(d) : rd32(d + 16), ft = rd32(d + 16);
        for (u32 k = 0;; k++) { u32 t = rd32(img + oft + 4 * k); if (!t) break; const char* nm = (const char*)(img + t + 2); ((u32*)(img + ft))[k] = (u32)lookup(nm); }
    }
}
static void* pe_export(const char* name) {
    u8* e = img + exp_rva; u32 n = rd32(e + 24), funcs = rd32(e + 28), names = rd32(e + 32), ords = rd32(e + 36);
    for (u32 i = 0; i < n; i++) if (!xstrcmp((char*)(img + rd32(img + names + 4 * i)), name)) return img + rd32(img + funcs + 4 * rd16(img + ords + 2 * i));
    put("missing export "); die(name); return NULL;
}

/* ------------------------------------------------------------------ file helpers */
static u8* slurp(const char* path, u32* len) {
    int fd = sys_open(path, 0, 0); if (fd < 0) die("cannot open input");
    u32 cap = 1u << 20, n = 0; u8* b = xmmap(0, cap, 3, 0x22);
    for (;;) {
        if (n == cap) { u8* nb = xmmap(0, cap * 2, 3, 0x22); if ((i32)nb < 0 && (i32)nb > -4096) die("mmap"); memcpy(nb, b, n); b = nb; cap *= 2; }
        i32 r = sys_read(fd, b + n, cap - n); if (r <= 0) break; n += (u32)r;
    }
    sys_close(fd); *len = n; return b;
}
static void spit(const char* path, const u8* p, u32 n) {
    int fd = sys_open(path, 0x241, 0644); if (fd < 0) die("cannot open output");
    while (n) { i32 r = sys_write(fd, p, n); if (r <= 0) die("write failed"); p += r; n -= (u32)r; }
    sys_close(fd);
}

/* ------------------------------------------------------------------ decoder client (the C# callbacks, restated) */
static const u8* in_buf; static u32 in_len, in_pos;
static u8* out_buf; static u32 out_cap, out_len;
static u32 n_frames, n_errors, err_codes[64], meta_sr, meta_ch, meta_bps; static u64 meta_total; static u8 meta_md5[16];
static u32 read_cap = 16384;   /* FLACDecoder.DEFAULT_MAX_BUFFER_SIZE (FLACDecoder.cs:21) */
typedef int (*fn_get_state)(void*);
static fn_get_state p_get_state; static u32 err_states[64];

static int use_ogg;      /* "decogg": the same client behind FLAC__stream_decoder_init_ogg_stream (the DLL carries libogg) */
/* ReadCallback, FLACDecoder.cs:325-363 */
ALIGNED static int dec_read(void* d, u8* buf, size_t* bytes, void* cd) {
    (void)d; (void)cd;
    u32 want = *bytes;
    if (want == 0) return 2;
    u32 length = want < read_cap ? want : read_cap;
    u32 left = in_len - in_pos, count = left < length ? left : length;
    memcpy(buf, in_buf + in_pos, count); in_pos += count;
    *bytes = count;
    /* The C# reports END_OF_STREAM together with the last, short read (FLACDecoder.cs:345-350).  libFLAC 1.2.1's Ogg layer
     * stops at that status without paging out what it has just been handed, so behind init_ogg_stream that convention loses
     * the frames of the last read; "decogg" follows libFLAC's own file callback instead: END_OF_STREAM only with 0 bytes. */
    if (use_ogg) return count == 0 ? 1 : 0;
    return count < length ? 1 : 0;
}
/* WriteCallback: FLACDecoder.cs:520-580 for 16-bit mono/stereo; FLACFileReader.cs:214-243 layout in general */
ALIGNED static int dec_write(void* d, const u8* frame, const i32* const* planes, void* cd) {
    (void)d; (void)cd;
    u32 bs = rd32(frame), ch = rd32(frame + 8), bps = rd32(frame + 16), B = (bps + 7) / 8;
    u32 need = bs * ch * B;
    if (out_len + need > out_cap) die("output buffer too small");
    u8* o = out_buf + out_len;
    for (u32 i = 0; i < bs; i++)
        for (u32 c = 0; c < ch; c++) { u32 v = (u32)planes[c][i]; for (u32 k = 0; k < B; k++) *o++ = (u8)(v >> (8 * k)); }
    out_len += need; n_frames++;
    return 0;
}
/* MetadataCallback, FLACDecoder.cs:431-473 (offsets = 12 + FieldOffset of LibFLACSharp.cs:295-319) */
ALIGNED static void dec_meta(void* d, const u8* m, void* cd) {
    (void)d; (void)cd;
    if (rd32(m) != 0) return;
    meta_sr = rd32(m + 32); meta_ch = rd32(m + 36); meta_bps = rd32(m + 40);
    meta_total = (u64)rd32(m + 48) | (u64)rd32(m + 52) << 32; memcpy(meta_md5, m + 56, 16);
}
/* ErrorCallback, FLACDecoder.cs:590-594 (the C# throws; here the status and state are recorded) */
ALIGNED static void dec_err(void* d, int status, void* cd) {
    (void)cd;
    if (n_errors < 64) { err_codes[n_errors] = (u32)status; err_states[n_errors] = (u32)p_get_state(d); }
    n_errors++;
}

typedef void* (*fn_new)(void); typedef int (*fn_i_p)(void*); typedef void (*fn_v_p)(void*);
typedef int (*fn_init)(void*, void*, void*, void*, void*, void*, void*, void*, void*, void*);

static int final_state;
static void decode_once(void) {
    fn_new dnew = (fn_new)pe_export("FLAC__stream_decoder_new");
    fn_init dinit = (fn_init)pe_export(use_ogg ? "FLAC__stream_decoder_init_ogg_stream" : "FLAC__stream_decoder_init_stream");
    fn_i_p meta = (fn_i_p)pe_export("FLAC__stream_decoder_process_until_end_of_metadata");
    fn_i_p single = (fn_i_p)pe_export("FLAC__stream_decoder_process_single");
    fn_i_p finish = (fn_i_p)pe_export("FLAC__stream_decoder_finish");
    fn_v_p del = (fn_v_p)pe_export("FLAC__stream_decoder_delete");
    p_get_state = (fn_get_state)pe_export("FLAC__stream_decoder_get_state");
    u32 mark = heap_off;
    in_pos = 0; out_len = 0; n_frames = 0; n_errors = 0;
    void* dec = dnew(); if (!dec) die("FLAC__stream_decoder_new failed");
    if (dinit(dec, (void*)dec_read, NULL, NULL, NULL, NULL, (void*)dec_write, (void*)dec_meta, (void*)dec_err, NULL) != 0) die("init_stream failed");
    if (!meta(dec)) { put("process_until_end_of_metadata failed state="); putu((u32)p_get_state(dec)); put("\n"); sys_exit(5); }
    while (p_get_state(dec) < 4) { if (!single(dec)) break; }
    final_state = p_get_state(dec);
    finish(dec); del(dec);
    heap_off = mark;   /* everything the decoder allocated is dead now */
}

/* ------------------------------------------------------------------ encoder client */
static u8* enc_buf; static u32 enc_cap, enc_len, enc_pos;
ALIGNED static int enc_write(void* e, const u8* buf, size_t bytes, u32 samples, u32 frame, void* cd) {
    (void)e; (void)samples; (void)frame; (void)cd;
    if (enc_pos + bytes > enc_cap) die("encoder output too large");
    memcpy(enc_buf + enc_pos, buf, bytes); enc_pos += bytes; if (enc_pos > enc_len) enc_len = enc_pos;
    return 0;
}
ALIGNED static int enc_seek(void* e, u64 off, void* cd) { (void)e; (void)cd; enc_pos = (u32)off; return 0; }
ALIGNED static int enc_tell(void* e, u64* off, void* cd) { (void)e; (void)cd; *off = enc_pos; return 0; }
typedef int (*fn_set_u)(void*, u32); typedef int (*fn_set_u64)(void*, u64);
typedef int (*fn_einit)(void*, void*, void*, void*, void*, void*); typedef int (*fn_proc)(void*, const i32*, u32);
static void set_u(void* enc, const char* name, u32 v) { if (!((fn_set_u)pe_export(name))(enc, v)) { put(name); die(" rejected"); } }

static int cmain(int argc, char** argv) {
    if (argc < 4) { put("usage: refflac dec|decogg|bench|enc <LibFlac.dll> ...\n"); return 2; }
    heap_cap = 768u << 20; heap = xmmap(0, heap_cap, 3, 0x22);
    if ((i32)heap < 0 && (i32)heap > -4096) die("cannot map heap");
    pe_load(argv[2]);
    if (!xstrcmp(argv[1], "dec") || !xstrcmp(argv[1], "bench") || !xstrcmp(argv[1], "decogg")) {
        int bench = argv[1][0] == 'b';
        use_ogg = !xstrcmp(argv[1], "decogg");
        in_buf = slurp(argv[3], &in_len);
        /* output capacity from STREAMINFO when present, else 64x input */
        out_cap = in_len * 8 + (64u << 20); if (out_cap > (1600u << 20) || out_cap < in_len) out_cap = 1600u << 20;
        out_buf = xmmap(0, out_cap, 3, 0x22); if ((i32)out_buf < 0 && (i32)out_buf > -4096) die("cannot map output");
        if (bench) {
            u32 iters = argc > 4 ? atou(argv[4]) : 3;
            for (u32 i = 0; i < iters; i++) { u64 t0 = now_ns(); decode_once(); u64 t1 = now_ns(); put("ms="); putu((u32)udiv64(t1 - t0, 1000)); put("us frames="); putu(n_frames); put(" bytes="); putu(out_len); put("\n"); }
            return 0;
        }
        decode_once();
        if (argc > 4) spit(argv[4], out_buf, out_len);
        put("sr="); putu(meta_sr); put(" ch="); putu(meta_ch); put(" bps="); putu(meta_bps); put(" total="); putu((u32)meta_total);
        put(" frames="); putu(n_frames); put(" bytes="); putu(out_len); put(" state="); putu((u32)final_state); put(" errors="); putu(n_errors);
        put(" si_md5="); for (int i = 0; i < 16; i++) puthex8(meta_md5[i]);
        put("\n");
        for (u32 i = 0; i < n_errors && i < 64; i++) { put("error["); putu(i); put("]="); putu(err_codes[i]); put(" state="); putu(err_states[i]); put("\n"); }
        return 0;
    }
    if (!xstrcmp(argv[1], "enc")) {
        if (argc < 14) die("enc needs: ch bps sr bs maxlpc minpo maxpo midside exhaustive in.pcm out.flac");
        u32 ch = atou(argv[3]), bps = atou(argv[4]), sr = atou(argv[5]), bs = atou(argv[6]), lpc = atou(argv[7]), minpo = atou(argv[8]),
            maxpo = atou(argv[9]), ms = atou(argv[10]), ex = atou(argv[11]);
        u32 plen; const u8* pcm = slurp(argv[12], &plen);
        u32 B = (bps + 7) / 8, n = plen / (B * ch);
        enc_cap = plen + (16u << 20); enc_buf = xmmap(0, enc_cap, 3, 0x22); enc_len = enc_pos = 0;
        void* enc = ((fn_new)pe_export("FLAC__stream_encoder_new"))(); if (!enc) die("encoder_new failed");
        set_u(enc, "FLAC__stream_encoder_set_streamable_subset", 0);
        set_u(enc, "FLAC__stream_encoder_set_channels", ch); set_u(enc, "FLAC__stream_encoder_set_bits_per_sample", bps);
        set_u(enc, "FLAC__stream_encoder_set_sample_rate", sr); set_u(enc, "FLAC__stream_encoder_set_compression_level", 5);
        set_u(enc, "FLAC__stream_encoder_set_blocksize", bs); set_u(enc, "FLAC__stream_encoder_set_max_lpc_order", lpc);
        set_u(enc, "FLAC__stream_encoder_set_min_residual_partition_order", minpo); set_u(enc, "FLAC__stream_encoder_set_max_residual_partition_order", maxpo);
        set_u(enc, "FLAC__stream_encoder_set_do_mid_side_stereo", ms); set_u(enc, "FLAC__stream_encoder_set_loose_mid_side_stereo", 0);
        set_u(enc, "FLAC__stream_encoder_set_do_exhaustive_model_search", ex);
        ((fn_set_u64)pe_export("FLAC__stream_encoder_set_total_samples_estimate"))(enc, n);
        if (((fn_einit)pe_export("FLAC__stream_encoder_init_stream"))(enc, (void*)enc_write, (void*)enc_seek, (void*)enc_tell, NULL, NULL) != 0) die("encoder init_stream failed");
        fn_proc proc = (fn_proc)pe_export("FLAC__stream_encoder_process_interleaved");
        static i32 chunk[4096 * 8];
        for (u32 i = 0; i < n;) {
            u32 c = n - i; if (c > 4096) c = 4096;
            for (u32 k = 0; k < c * ch; k++) {
                const u8* q = pcm + (i * ch + k) * B; u32 v = 0;
                for (u32 b = 0; b < B; b++) v |= (u32)q[b] << (8 * b);
                chunk[k] = (i32)(v << (32 - 8 * B)) >> (32 - 8 * B);
            }
            if (!proc(enc, chunk, c)) die("process_interleaved failed");
            i += c;
        }
        if (!((fn_i_p)pe_export("FLAC__stream_encoder_finish"))(enc)) die("encoder finish failed");
        ((fn_v_p)pe_export("FLAC__stream_encoder_delete"))(enc);
        spit(argv[13], enc_buf, enc_len);
        put("encoded samples="); putu(n); put(" bytes="); putu(enc_len); put("\n");
        return 0;
    }
    die("unknown mode");
    return 2;
}

/* entry: pass the original stack pointer (argc, argv) to C */
void start_c(u32* sp) { int argc = (int)sp[0]; char** argv = (char**)(sp + 1); sys_exit(cmain(argc, argv)); }
__asm__(".text\n.globl _start\n_start:\n mov %esp,%eax\n and $-16,%esp\n sub $12,%esp\n push %eax\n call start_c\n hlt\n");
