/*
 * orc_shim.c -- stand-ins for the external programs the reference shells out to, backed by the
 * CPU oracle (dandd_oracle.c).  TEST INFRASTRUCTURE ONLY (see the oracle header).
 *
 * Dispatches on basename(argv[0]) (install as symlinks named dashing / kmc / kmc_tools) or, if
 * argv[0] is "orc_shim", on argv[1].  Implements exactly the argv / stdout surface DandD uses
 * (SURVEY.md section 8b):
 *   dashing sketch [--no-canon] -k<K> -S <p> --prefix <dir> <fasta>   lib/sketch_classes.py:351-366
 *   dashing union -z -o <out> <in...>                                 lib/sketch_classes.py:368-373
 *   dashing card --presketched <paths...>                             lib/sketch_classes.py:306-321
 *   dashing hll -k <K> -S <p> <fasta...>                              helpers/allpairs.py:32-35
 *   kmc -hp [-tN] -ci1 -cs2 -k<K> [-b] -fm <fasta> <out> <tmp>        lib/sketch_classes.py:434-449
 *   kmc_tools -hp info <db>                                           lib/sketch_classes.py:389-399
 *   kmc_tools -hp [-tN] complex /dev/stdin                            lib/sketch_classes.py:451-465
 * Sketch files follow SURVEY.md A.7 (gz stream: u32[4] flags, u32 p, f64 cached value, 2^p u8).
 * The KMC database files are a private format (sorted distinct u64 k-mers); DandD never opens
 * them, it only checks that .kmc_pre/.kmc_suf exist and are non-empty.
 */
#define _GNU_SOURCE
#include <libgen.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

size_t orc_fasta_symbols(const uint8_t *buf, size_t n, uint8_t *out);
void orc_hll_sketch(const uint8_t *sym, size_t n, int k, int p, int canon, uint8_t *regs);
double orc_card(const uint8_t *regs, int p);
size_t orc_kmers(const uint8_t *sym, size_t n, int k, int canon, uint64_t *out, size_t cap);

static void die(const char *msg, const char *arg) {
    fprintf(stderr, "orc_shim: %s %s\n", msg, arg ? arg : "");
    exit(2);
}

/* Read a whole (possibly gzip-compressed) file. */
static uint8_t *slurp(const char *path, size_t *n_out) {
    gzFile f = gzopen(path, "rb");
    if (!f) die("cannot open", path);
    size_t cap = 1u << 20, n = 0;
    uint8_t *buf = (uint8_t *)malloc(cap);
    for (;;) {
        if (n == cap) { cap *= 2; buf = (uint8_t *)realloc(buf, cap); }
        int got = gzread(f, buf + n, (unsigned)((cap - n) > (1u << 30) ? (1u << 30) : (cap - n)));
        if (got < 0) die("read error", path);
        if (got == 0) break;
        n += (size_t)got;
    }
    gzclose(f);
    *n_out = n;
    return buf;
}

static uint8_t *fasta_symbols(const char *path, size_t *nsym) {
    size_t n;
    uint8_t *txt = slurp(path, &n);
    uint8_t *sym = (uint8_t *)malloc(n ? n : 1);
    *nsym = orc_fasta_symbols(txt, n, sym);
    free(txt);
    return sym;
}

/* ---- .hll files (A.7) ---- */
static void hll_write(const char *path, const uint8_t *regs, int p, int compress) {
    gzFile f = gzopen(path, compress ? "wb1" : "wbT");
    if (!f) die("cannot write", path);
    uint32_t flags[4] = {0, 2, 2, 3}; /* the Ertl MLE under both recalled readings of the header: dandd_b200/hllfile.py */
    uint32_t np = (uint32_t)p;
    double value = 0.0;
    gzwrite(f, flags, sizeof flags);
    gzwrite(f, &np, sizeof np);
    gzwrite(f, &value, sizeof value);
    gzwrite(f, regs, 1u << p);
    gzclose(f);
}

static uint8_t *hll_read(const char *path, int *p_out) {
    size_t n;
    uint8_t *raw = slurp(path, &n);
    if (n < 28) die("short sketch", path);
    uint32_t np;
    memcpy(&np, raw + 16, 4);
    if (np > 32 || n != 28 + ((size_t)1 << np)) die("bad sketch", path);
    uint8_t *regs = (uint8_t *)malloc((size_t)1 << np);
    memcpy(regs, raw + 28, (size_t)1 << np);
    free(raw);
    *p_out = (int)np;
    return regs;
}

static int starts(const char *s, const char *pre) { return strncmp(s, pre, strlen(pre)) == 0; }

static int dashing_main(int argc, char **argv) {
    if (argc < 2) die("dashing: missing subcommand", NULL);
    const char *sub = argv[1];
    if (!strcmp(sub, "sketch") || !strcmp(sub, "hll")) {
        int k = 31, p = 10, canon = 1;
        const char *prefix = ".";
        int nfiles = 0;
        char **files = (char **)calloc((size_t)argc, sizeof(char *));
        for (int i = 2; i < argc; ++i) {
            char *a = argv[i];
            if (!strcmp(a, "--no-canon")) canon = 0;
            else if (!strcmp(a, "--prefix")) prefix = argv[++i];
            else if (!strcmp(a, "-k")) k = atoi(argv[++i]);
            else if (starts(a, "-k")) k = atoi(a + 2);
            else if (!strcmp(a, "-S")) p = atoi(argv[++i]);
            else if (starts(a, "-S")) p = atoi(a + 2);
            else if (starts(a, "-p")) { if (!a[2]) ++i; }
            else if (a[0] == '-') die("dashing: unsupported flag", a);
            else files[nfiles++] = a;
        }
        if (k < 1 || k > 32) die("dashing: k out of range", NULL);
        if (!strcmp(sub, "hll")) { /* all files into ONE sketch; last token = estimate */
            uint8_t *regs = (uint8_t *)calloc((size_t)1 << p, 1);
            for (int f = 0; f < nfiles; ++f) {
                size_t ns; uint8_t *sym = fasta_symbols(files[f], &ns);
                orc_hll_sketch(sym, ns, k, p, canon, regs);
                free(sym);
            }
            printf("Estimated number of unique exact matches: %lf\n", orc_card(regs, p));
            return 0;
        }
        for (int f = 0; f < nfiles; ++f) {
            size_t ns; uint8_t *sym = fasta_symbols(files[f], &ns);
            uint8_t *regs = (uint8_t *)calloc((size_t)1 << p, 1);
            orc_hll_sketch(sym, ns, k, p, canon, regs);
            char *dup = strdup(files[f]);
            char out[4096];
            snprintf(out, sizeof out, "%s/%s.w.%d.spacing.%d.hll", prefix, basename(dup), k, p);
            hll_write(out, regs, p, 1);
            free(dup); free(regs); free(sym);
        }
        return 0;
    }
    if (!strcmp(sub, "union")) {
        const char *out = NULL; int compress = 0, nin = 0;
        char **in = (char **)calloc((size_t)argc, sizeof(char *));
        for (int i = 2; i < argc; ++i) {
            char *a = argv[i];
            if (!strcmp(a, "-z")) compress = 1;
            else if (!strcmp(a, "-o")) out = argv[++i];
            else if (starts(a, "-p")) { if (!a[2]) ++i; }
            else if (a[0] == '-') die("dashing union: unsupported flag", a);
            else in[nin++] = a;
        }
        if (!out || nin == 0) die("dashing union: need -o and inputs", NULL);
        int p = -1; uint8_t *acc = NULL;
        for (int j = 0; j < nin; ++j) {
            int pj; uint8_t *r = hll_read(in[j], &pj);
            if (p < 0) { p = pj; acc = (uint8_t *)calloc((size_t)1 << p, 1); }
            if (pj != p) die("dashing union: mismatched sketch sizes", in[j]);
            for (size_t i = 0; i < ((size_t)1 << p); ++i) if (r[i] > acc[i]) acc[i] = r[i];
            free(r);
        }
        hll_write(out, acc, p, compress);
        return 0;
    }
    if (!strcmp(sub, "card")) {
        printf("#Path\tSize (est.)\n");
        for (int i = 2; i < argc; ++i) {
            if (argv[i][0] == '-') continue; /* --presketched, -pN */
            int p; uint8_t *r = hll_read(argv[i], &p);
            printf("%s\t%lf\n", argv[i], orc_card(r, p));
            free(r);
        }
        return 0;
    }
    die("dashing: unsupported subcommand", sub);
    return 2;
}

/* ---- KMC stand-in: database = sorted distinct u64 k-mers ---- */
/* A shim database is the sorted list of distinct k-mers as fixed-width records: 8 bytes (the 2-bit value as a
 * uint64) for k <= 32, k bytes (one symbol per byte, what KMC's k <= 256 needs) above.  Private to this shim. */
static size_t g_width = 8;
static int cmp_rec(const void *a, const void *b) {
    if (g_width == 8) {
        uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
        return x < y ? -1 : x > y;
    }
    return memcmp(a, b, g_width);
}

static void kmcdb_write(const char *base, const uint8_t *recs, uint64_t n, size_t width, int k, int canon) {
    char path[4096];
    snprintf(path, sizeof path, "%s.kmc_pre", base);
    FILE *f = fopen(path, "wb");
    if (!f) die("cannot write", path);
    uint64_t hdr[5] = {0x434d4b43524fULL /* "ORCKMC" */, (uint64_t)k, (uint64_t)canon, n, (uint64_t)width};
    fwrite(hdr, sizeof hdr, 1, f);
    fclose(f);
    snprintf(path, sizeof path, "%s.kmc_suf", base);
    f = fopen(path, "wb");
    if (!f) die("cannot write", path);
    uint64_t magic = 0x46555343524fULL;
    fwrite(&magic, sizeof magic, 1, f); /* never empty, even for 0 k-mers */
    fwrite(recs, width, n, f);
    fclose(f);
}

static uint8_t *kmcdb_read(const char *base, uint64_t *n_out, size_t *width_out) {
    char path[4096];
    uint64_t hdr[5];
    snprintf(path, sizeof path, "%s.kmc_pre", base);
    FILE *f = fopen(path, "rb");
    if (!f || fread(hdr, sizeof hdr, 1, f) != 1) die("cannot read db", path);
    fclose(f);
    snprintf(path, sizeof path, "%s.kmc_suf", base);
    f = fopen(path, "rb");
    if (!f) die("cannot read db", path);
    uint64_t magic, n = hdr[3];
    size_t width = (size_t)hdr[4];
    uint8_t *a = (uint8_t *)malloc((n ? n : 1) * width);
    if (fread(&magic, sizeof magic, 1, f) != 1 || fread(a, width, n, f) != n) die("short db", path);
    fclose(f);
    *n_out = n;
    *width_out = width;
    return a;
}

static uint64_t uniq(uint8_t *a, uint64_t n, size_t width) {
    if (!n) return 0;
    g_width = width;
    qsort(a, n, width, cmp_rec);
    uint64_t o = 1;
    for (uint64_t i = 1; i < n; ++i)
        if (memcmp(a + i * width, a + (o - 1) * width, width)) memmove(a + (o++) * width, a + i * width, width);
    return o;
}

/* k > 32: every valid window of the symbol stream as a k-byte record (canonical = the smaller of the window and
 * its reverse complement, compared as strings -- the same order as comparing the 2-bit integers). */
static uint8_t *long_kmers(const uint8_t *sym, size_t ns, int k, int canon, uint64_t *cnt_out) {
    uint64_t cnt = 0;
    size_t run = 0;
    for (size_t i = 0; i < ns; ++i) { run = sym[i] > 3 ? 0 : run + 1; if (run >= (size_t)k) ++cnt; }
    uint8_t *out = (uint8_t *)malloc((cnt ? cnt : 1) * (size_t)k), *rc = (uint8_t *)malloc((size_t)k);
    uint64_t o = 0;
    run = 0;
    for (size_t i = 0; i < ns; ++i) {
        run = sym[i] > 3 ? 0 : run + 1;
        if (run < (size_t)k) continue;
        const uint8_t *w = sym + i + 1 - k;
        const uint8_t *pick = w;
        if (canon) {
            for (int j = 0; j < k; ++j) rc[j] = (uint8_t)(3 - w[k - 1 - j]);
            if (memcmp(rc, w, (size_t)k) < 0) pick = rc;
        }
        memcpy(out + o * (size_t)k, pick, (size_t)k);
        ++o;
    }
    free(rc);
    *cnt_out = cnt;
    return out;
}

static int kmc_main(int argc, char **argv) {
    int k = 25, canon = 1;
    char *pos[3]; int npos = 0;
    for (int i = 1; i < argc; ++i) {
        char *a = argv[i];
        if (starts(a, "-k")) k = atoi(a + 2);
        else if (!strcmp(a, "-b")) canon = 0;
        else if (a[0] == '-') continue; /* -hp -tN -ci1 -cs2 -fm */
        else if (npos < 3) pos[npos++] = a;
    }
    if (npos < 2) die("kmc: need <input> <output> <tmp>", NULL);
    if (k < 1 || k > 256) die("kmc: k-mer length must be in [1, 256]", NULL);
    size_t ns; uint8_t *sym = fasta_symbols(pos[0], &ns);
    if (k <= 32) {
        size_t cnt = orc_kmers(sym, ns, k, canon, NULL, 0);
        uint64_t *a = (uint64_t *)malloc((cnt ? cnt : 1) * sizeof(uint64_t));
        orc_kmers(sym, ns, k, canon, a, cnt);
        uint64_t n = uniq((uint8_t *)a, cnt, 8);
        kmcdb_write(pos[1], (uint8_t *)a, n, 8, k, canon);
    } else {
        uint64_t cnt; uint8_t *a = long_kmers(sym, ns, k, canon, &cnt);
        uint64_t n = uniq(a, cnt, (size_t)k);
        kmcdb_write(pos[1], a, n, (size_t)k, k, canon);
    }
    return 0;
}

static int kmc_tools_main(int argc, char **argv) {
    int i = 1;
    while (i < argc && argv[i][0] == '-') ++i;
    if (i >= argc) die("kmc_tools: missing operation", NULL);
    if (!strcmp(argv[i], "info")) {
        uint64_t n; size_t width; uint8_t *a = kmcdb_read(argv[i + 1], &n, &width);
        free(a);
        printf("k                 :  0\ntotal k-mers      :  %llu\n", (unsigned long long)n);
        return 0;
    }
    if (!strcmp(argv[i], "complex")) {
        /* INPUT:\n inputN = <path> -ci1 \n ... OUTPUT:\n <out> = input1 + input2 ...  (all unions) */
        FILE *f = fopen(argv[i + 1], "r");
        if (!f) die("kmc_tools: cannot open", argv[i + 1]);
        char line[65536], out[4096] = "";
        uint8_t *acc = NULL;
        uint64_t nacc = 0;
        size_t width = 8;
        int in_output = 0;
        while (fgets(line, sizeof line, f)) {
            char name[4096], path[4096];
            if (strstr(line, "INPUT:")) { in_output = 0; continue; }
            if (strstr(line, "OUTPUT:")) { in_output = 1; continue; }
            if (!in_output && sscanf(line, " %4095s = %4095s", name, path) == 2) {
                uint64_t n; size_t w; uint8_t *a = kmcdb_read(path, &n, &w);
                if (nacc && w != width) die("kmc_tools complex: inputs of different k", path);
                width = w;
                acc = (uint8_t *)realloc(acc, (nacc + n + 1) * width);
                memcpy(acc + nacc * width, a, n * width);
                nacc += n; free(a);
            } else if (in_output && sscanf(line, " %4095s =", out) == 1) {
                break;
            }
        }
        fclose(f);
        if (!out[0]) die("kmc_tools complex: no OUTPUT", NULL);
        uint64_t n = uniq(acc, nacc, width);
        kmcdb_write(out, acc, n, width, 0, 1);
        return 0;
    }
    die("kmc_tools: unsupported operation", argv[i]);
    return 2;
}

int main(int argc, char **argv) {
    char *dup = strdup(argv[0]);
    const char *me = basename(dup);
    if (!strcmp(me, "orc_shim")) {
        if (argc < 2) die("usage: orc_shim {dashing|kmc|kmc_tools} ...", NULL);
        me = argv[1]; ++argv; --argc;
    }
    if (!strcmp(me, "dashing")) return dashing_main(argc, argv);
    if (!strcmp(me, "kmc")) return kmc_main(argc, argv);
    if (!strcmp(me, "kmc_tools")) return kmc_tools_main(argc, argv);
    die("unknown personality", me);
    return 2;
}
