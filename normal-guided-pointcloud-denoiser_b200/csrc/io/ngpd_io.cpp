// libngpd_io.so: threaded readers / writer for the files of the point-cloud path (include/ngpd_io.h).
// Host code only (g++, no CUDA).  Replaces igl.read_obj (Object.py:80), the Python line loops of loadXYZ / saveObj
// (Object.py:58-69, 92-117) and Open3D's ASCII .ply body (Object.py:127) for clouds of 10^7 - 10^8 points.
//
// Layout of every reader: map the file, cut [begin, end) into one piece per thread at line ends, parse the pieces
// independently into per-piece vectors, stitch them in file order.  The only cross-piece dependency of an OBJ file are
// RELATIVE ids (`f -1 -2 -3` counts back from the records seen so far): a piece stores them relative to its own start
// (marked) and the stitch adds the number of records in the pieces before it.
#include "../../../include/ngpd_io.h"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

struct Mapped {
    const char* p = nullptr;
    size_t size = 0;
    int fd = -1;
    ~Mapped() {
        if (p && size) munmap(const_cast<char*>(p), size);
        if (fd >= 0) close(fd);
    }
    int open_file(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return fail(NGPD_IO_ERR_OPEN, "cannot open %s: %s", path, strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return fail(NGPD_IO_ERR_OPEN, "%s is not a regular file", path);
        size = (size_t)st.st_size;
        if (size == 0) return 0;
        void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { size = 0; return fail(NGPD_IO_ERR_OPEN, "cannot map %s: %s", path, strerror(errno)); }
        madvise(m, size, MADV_SEQUENTIAL);
        p = (const char*)m;
        return 0;
    }
};

int pick_threads(int threads, size_t bytes) {
    if (threads <= 0) threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    // a piece below ~1 MB is not worth a thread
    size_t useful = bytes / (1u << 20) + 1;
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, useful));
}

// piece boundaries at line starts: cut[i] .. cut[i + 1]
std::vector<size_t> cut_at_lines(const char* p, size_t begin, size_t end, int pieces) {
    std::vector<size_t> cut(pieces + 1, end);
    cut[0] = begin;
    for (int i = 1; i < pieces; ++i) {
        size_t guess = begin + (end - begin) / pieces * i;
        guess = std::max(guess, cut[i - 1]);
        const void* nl = guess < end ? memchr(p + guess, '\n', end - guess) : nullptr;
        cut[i] = nl ? (size_t)((const char*)nl - p) + 1 : end;
    }
    return cut;
}

template <class F>
void run_pieces(int pieces, F&& body) {
    if (pieces == 1) { body(0); return; }
    std::vector<std::thread> pool;
    pool.reserve(pieces);
    for (int i = 0; i < pieces; ++i) pool.emplace_back([&body, i] { body(i); });
    for (auto& t : pool) t.join();
}

inline bool blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

// one number; advances q.  false = no number here.
inline bool parse_double(const char*& q, const char* e, double& out) {
    while (q < e && blank(*q)) ++q;
    if (q < e && *q == '+') ++q;
    auto r = std::from_chars(q, e, out);
    if (r.ec != std::errc() && r.ec != std::errc::result_out_of_range) return false;
    q = r.ptr;
    return true;
}
inline bool parse_int(const char*& q, const char* e, int64_t& out) {
    if (q < e && *q == '+') ++q;
    auto r = std::from_chars(q, e, out);
    if (r.ec != std::errc()) return false;
    q = r.ptr;
    return true;
}

constexpr int64_t REL_MARK = (int64_t)1 << 62;   // id stored relative to the start of its piece
constexpr int64_t REL_TEST = (int64_t)1 << 61;

struct ObjPiece {
    std::vector<double> v, vn;
    std::vector<int64_t> f, fn;
    size_t bad_at = SIZE_MAX;
};

// id of a face corner: positive = 1-based absolute, negative = relative to the `seen` records before this line
inline int64_t corner_id(int64_t raw, int64_t seen) { return raw > 0 ? raw - 1 : (seen + raw) + REL_MARK; }

void parse_obj_piece(const char* p, size_t b, size_t e, ObjPiece& out) {
    const char* q = p + b;
    const char* end = p + e;
    std::vector<int64_t> cv, cn;   // corners of the current face
    std::vector<char> hn;          // corner carries a normal id
    while (q < end) {
        const char* eol = (const char*)memchr(q, '\n', end - q);
        if (!eol) eol = end;
        const char* s = q;
        while (s < eol && blank(*s)) ++s;
        if (eol - s >= 2 && s[0] == 'v' && blank(s[1])) {
            const char* t = s + 1;
            double x[3];
            if (!parse_double(t, eol, x[0]) || !parse_double(t, eol, x[1]) || !parse_double(t, eol, x[2])) { out.bad_at = (size_t)(s - p); return; }
            out.v.insert(out.v.end(), x, x + 3);
        } else if (eol - s >= 3 && s[0] == 'v' && s[1] == 'n' && blank(s[2])) {
            const char* t = s + 2;
            double x[3];
            if (!parse_double(t, eol, x[0]) || !parse_double(t, eol, x[1]) || !parse_double(t, eol, x[2])) { out.bad_at = (size_t)(s - p); return; }
            out.vn.insert(out.vn.end(), x, x + 3);
        } else if (eol - s >= 2 && s[0] == 'f' && blank(s[1])) {
            const char* t = s + 1;
            cv.clear(); cn.clear(); hn.clear();
            const int64_t seen_v = (int64_t)(out.v.size() / 3), seen_n = (int64_t)(out.vn.size() / 3);
            for (;;) {
                while (t < eol && blank(*t)) ++t;
                if (t >= eol) break;
                int64_t a = 0;
                if (!parse_int(t, eol, a) || a == 0) { out.bad_at = (size_t)(s - p); return; }
                cv.push_back(corner_id(a, seen_v));
                bool has_n = false;
                if (t < eol && *t == '/') {
                    ++t;
                    int64_t tc;
                    if (t < eol && *t != '/' && !blank(*t)) { if (!parse_int(t, eol, tc)) { out.bad_at = (size_t)(s - p); return; } }
                    if (t < eol && *t == '/') {
                        ++t;
                        int64_t nn = 0;
                        if (t < eol && !blank(*t)) {
                            if (!parse_int(t, eol, nn) || nn == 0) { out.bad_at = (size_t)(s - p); return; }
                            cn.push_back(corner_id(nn, seen_n));
                            has_n = true;
                        }
                    }
                }
                if (!has_n) cn.push_back(0);
                hn.push_back(has_n);
            }
            // fan triangulation; a triangle carries normal ids only when all three corners do
            for (size_t a = 1; a + 1 < cv.size(); ++a) {
                const int64_t tri[3] = {cv[0], cv[a], cv[a + 1]};
                out.f.insert(out.f.end(), tri, tri + 3);
                if (hn[0] && hn[a] && hn[a + 1]) {
                    const int64_t trn[3] = {cn[0], cn[a], cn[a + 1]};
                    out.fn.insert(out.fn.end(), trn, trn + 3);
                }
            }
        }
        q = eol < end ? eol + 1 : end;
    }
}

struct TablePiece {
    std::vector<double> rows;
    size_t bad_at = SIZE_MAX;
};

inline bool skipped_line(const char* s, const char* eol) {
    while (s < eol && blank(*s)) ++s;
    return s >= eol || *s == '#';
}

void parse_table_piece(const char* p, size_t b, size_t e, int cols, TablePiece& out) {
    const char* q = p + b;
    const char* end = p + e;
    double x[64];
    while (q < end) {
        const char* eol = (const char*)memchr(q, '\n', end - q);
        if (!eol) eol = end;
        if (!skipped_line(q, eol)) {
            const char* t = q;
            for (int c = 0; c < cols; ++c)
                if (!parse_double(t, eol, x[c])) { out.bad_at = (size_t)(q - p); return; }
            out.rows.insert(out.rows.end(), x, x + cols);
        }
        q = eol < end ? eol + 1 : end;
    }
}

// ---- Python's repr(float) ------------------------------------------------------------------------------------------
// shortest digits that round-trip the double, laid out like CPython's float_repr_style 'short': exponent form when the
// decimal point would sit more than 16 digits right or more than 3 zeros left of the first digit, else positional with at
// least one digit after the point.
inline char* py_repr(double x, char* o) {
    if (std::isnan(x)) { memcpy(o, "nan", 3); return o + 3; }
    if (std::isinf(x)) { if (x < 0) *o++ = '-'; memcpy(o, "inf", 3); return o + 3; }
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);
    const char* s = buf;
    if (*s == '-') { *o++ = '-'; ++s; }
    char digits[24];
    int nd = 0;
    while (s < r.ptr && *s != 'e') { if (*s != '.') digits[nd++] = *s; ++s; }
    int ex = 0;
    if (s < r.ptr && *s == 'e') {
        ++s;
        bool neg = *s == '-';
        if (*s == '-' || *s == '+') ++s;
        while (s < r.ptr) ex = ex * 10 + (*s++ - '0');
        if (neg) ex = -ex;
    }
    const int decpt = ex + 1;                        // position of the decimal point relative to the first digit
    if (decpt > 16 || decpt < -3) {
        *o++ = digits[0];
        if (nd > 1) { *o++ = '.'; memcpy(o, digits + 1, nd - 1); o += nd - 1; }
        *o++ = 'e';
        int e10 = decpt - 1;
        *o++ = e10 < 0 ? '-' : '+';
        if (e10 < 0) e10 = -e10;
        if (e10 >= 100) { *o++ = (char)('0' + e10 / 100); e10 %= 100; *o++ = (char)('0' + e10 / 10); *o++ = (char)('0' + e10 % 10); }
        else { *o++ = (char)('0' + e10 / 10); *o++ = (char)('0' + e10 % 10); }
        return o;
    }
    if (decpt <= 0) {
        *o++ = '0'; *o++ = '.';
        for (int i = 0; i < -decpt; ++i) *o++ = '0';
        memcpy(o, digits, nd);
        return o + nd;
    }
    if (decpt >= nd) {
        memcpy(o, digits, nd); o += nd;
        for (int i = nd; i < decpt; ++i) *o++ = '0';
        *o++ = '.'; *o++ = '0';
        return o;
    }
    memcpy(o, digits, decpt); o += decpt;
    *o++ = '.';
    memcpy(o, digits + decpt, nd - decpt);
    return o + (nd - decpt);
}

bool write_all(int fd, const char* p, size_t n) {
    while (n) {
        ssize_t w = ::write(fd, p, n);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        p += w; n -= (size_t)w;
    }
    return true;
}

// "<tag> x y z\n" for rows [0, count) of a [count, 3] float array, in blocks formatted by the pool and written in order
template <class T>
int write_records(int fd, const char* tag, const T* a, int64_t count, int threads) {
    const size_t taglen = strlen(tag);
    const int64_t block = 1 << 18;                                   // rows per thread and round (~15 MB of text)
    std::vector<std::string> text(threads);
    for (int64_t base = 0; base < count; base += block * threads) {
        const int used = (int)std::min<int64_t>(threads, (count - base + block - 1) / block);
        run_pieces(used, [&](int t) {
            const int64_t i0 = base + block * t, i1 = std::min(count, i0 + block);
            std::string& s = text[t];
            s.resize((size_t)(i1 - i0) * (taglen + 3 * 26 + 4));
            char* o = &s[0];
            for (int64_t i = i0; i < i1; ++i) {
                memcpy(o, tag, taglen); o += taglen;
                for (int c = 0; c < 3; ++c) { *o++ = ' '; o = py_repr((double)a[3 * i + c], o); }
                *o++ = '\n';
            }
            s.resize((size_t)(o - &s[0]));
        });
        for (int t = 0; t < used; ++t)
            if (!write_all(fd, text[t].data(), text[t].size())) return fail(NGPD_IO_ERR_OPEN, "write failed: %s", strerror(errno));
    }
    return 0;
}

}  // namespace

struct ngpd_io_arrays {
    std::vector<double> v, vn, table;
    std::vector<int64_t> f, fn;
    int cols = 3;
};

#define NGPD_IO_API extern "C" __attribute__((visibility("default")))

NGPD_IO_API const char* ngpd_io_last_error(void) { return g_error.c_str(); }

NGPD_IO_API int ngpd_io_read_obj(const char* path, int threads, ngpd_io_arrays_t** out) {
    if (!path || !out) return fail(NGPD_IO_ERR_ARG, "ngpd_io_read_obj: NULL argument");
    *out = nullptr;
    Mapped m;
    if (int rc = m.open_file(path)) return rc;
    const int pieces = pick_threads(threads, m.size);
    std::vector<size_t> cut = cut_at_lines(m.p, 0, m.size, pieces);
    std::vector<ObjPiece> part(pieces);
    run_pieces(pieces, [&](int i) { parse_obj_piece(m.p, cut[i], cut[i + 1], part[i]); });
    for (int i = 0; i < pieces; ++i)
        if (part[i].bad_at != SIZE_MAX) return fail(NGPD_IO_ERR_PARSE, "%s: malformed record at byte %zu", path, part[i].bad_at);
    auto* A = new ngpd_io_arrays();
    size_t nv = 0, nn = 0, nf = 0, nfn = 0;
    for (auto& q : part) { nv += q.v.size(); nn += q.vn.size(); nf += q.f.size(); nfn += q.fn.size(); }
    A->v.resize(nv); A->vn.resize(nn); A->f.resize(nf); A->fn.resize(nfn);
    std::vector<size_t> ov(pieces + 1, 0), on(pieces + 1, 0), of(pieces + 1, 0), ofn(pieces + 1, 0);
    for (int i = 0; i < pieces; ++i) {
        ov[i + 1] = ov[i] + part[i].v.size(); on[i + 1] = on[i] + part[i].vn.size();
        of[i + 1] = of[i] + part[i].f.size(); ofn[i + 1] = ofn[i] + part[i].fn.size();
    }
    std::atomic<int64_t> bad{-1};
    run_pieces(pieces, [&](int i) {
        if (!part[i].v.empty()) memcpy(A->v.data() + ov[i], part[i].v.data(), part[i].v.size() * sizeof(double));
        if (!part[i].vn.empty()) memcpy(A->vn.data() + on[i], part[i].vn.data(), part[i].vn.size() * sizeof(double));
        const int64_t base_v = (int64_t)(ov[i] / 3), base_n = (int64_t)(on[i] / 3);
        for (size_t k = 0; k < part[i].f.size(); ++k) {
            int64_t x = part[i].f[k];
            if (x >= REL_TEST) x = x - REL_MARK + base_v;
            if (x < 0 || x >= (int64_t)(nv / 3)) bad.store((int64_t)k);
            A->f[of[i] + k] = x;
        }
        for (size_t k = 0; k < part[i].fn.size(); ++k) {
            int64_t x = part[i].fn[k];
            if (x >= REL_TEST) x = x - REL_MARK + base_n;
            if (x < 0 || x >= (int64_t)(nn / 3)) bad.store((int64_t)k);
            A->fn[ofn[i] + k] = x;
        }
    });
    if (bad.load() >= 0) { delete A; return fail(NGPD_IO_ERR_PARSE, "%s: a face refers to a vertex or normal that does not exist", path); }
    *out = A;
    return 0;
}

NGPD_IO_API int ngpd_io_read_table(const char* path, int64_t offset, int64_t rows, int cols, int threads, ngpd_io_arrays_t** out) {
    if (!path || !out || cols < 1 || cols > 64 || offset < 0) return fail(NGPD_IO_ERR_ARG, "ngpd_io_read_table: bad argument (1 <= cols <= 64)");
    *out = nullptr;
    Mapped m;
    if (int rc = m.open_file(path)) return rc;
    if ((size_t)offset > m.size) return fail(NGPD_IO_ERR_ARG, "%s: offset %lld is past the end of the file", path, (long long)offset);
    size_t end = m.size;
    if (rows >= 0) {
        // the first `rows` data lines end here
        const char* q = m.p + offset;
        const char* fe = m.p + m.size;
        int64_t left = rows;
        while (left > 0 && q < fe) {
            const char* eol = (const char*)memchr(q, '\n', fe - q);
            if (!eol) eol = fe;
            if (!skipped_line(q, eol)) --left;
            q = eol < fe ? eol + 1 : fe;
        }
        if (left > 0) return fail(NGPD_IO_ERR_PARSE, "%s: %lld rows expected, the file ends %lld rows early", path, (long long)rows, (long long)left);
        end = (size_t)(q - m.p);
    }
    const int pieces = pick_threads(threads, end - (size_t)offset);
    std::vector<size_t> cut = cut_at_lines(m.p, (size_t)offset, end, pieces);
    std::vector<TablePiece> part(pieces);
    run_pieces(pieces, [&](int i) { parse_table_piece(m.p, cut[i], cut[i + 1], cols, part[i]); });
    for (int i = 0; i < pieces; ++i)
        if (part[i].bad_at != SIZE_MAX) return fail(NGPD_IO_ERR_PARSE, "%s: fewer than %d numbers in the line at byte %zu", path, cols, part[i].bad_at);
    auto* A = new ngpd_io_arrays();
    A->cols = cols;
    size_t total = 0;
    std::vector<size_t> off(pieces + 1, 0);
    for (int i = 0; i < pieces; ++i) { off[i + 1] = off[i] + part[i].rows.size(); }
    total = off[pieces];
    A->table.resize(total);
    run_pieces(pieces, [&](int i) {
        if (!part[i].rows.empty()) memcpy(A->table.data() + off[i], part[i].rows.data(), part[i].rows.size() * sizeof(double));
    });
    *out = A;
    return 0;
}

NGPD_IO_API int64_t ngpd_io_count(const ngpd_io_arrays_t* a, int which) {
    if (!a) return 0;
    switch (which) {
        case NGPD_IO_VERTICES: return (int64_t)(a->v.size() / 3);
        case NGPD_IO_NORMALS: return (int64_t)(a->vn.size() / 3);
        case NGPD_IO_FACES: return (int64_t)(a->f.size() / 3);
        case NGPD_IO_FACE_NORMALS: return (int64_t)(a->fn.size() / 3);
        case NGPD_IO_TABLE: return (int64_t)(a->table.size() / (size_t)a->cols);
    }
    return 0;
}

NGPD_IO_API const void* ngpd_io_data(const ngpd_io_arrays_t* a, int which) {
    if (!a) return nullptr;
    switch (which) {
        case NGPD_IO_VERTICES: return a->v.data();
        case NGPD_IO_NORMALS: return a->vn.data();
        case NGPD_IO_FACES: return a->f.data();
        case NGPD_IO_FACE_NORMALS: return a->fn.data();
        case NGPD_IO_TABLE: return a->table.data();
    }
    return nullptr;
}

NGPD_IO_API void ngpd_io_free(ngpd_io_arrays_t* a) { delete a; }

template <class T>
int write_obj(const char* path, const T* v, const T* n, int64_t count, int exclusive, int threads) {
    if (!path || (!v && count > 0) || count < 0) return fail(NGPD_IO_ERR_ARG, "ngpd_io_write_obj: bad argument");
    int fd = ::open(path, O_WRONLY | O_CREAT | (exclusive ? O_EXCL : O_TRUNC), 0644);
    if (fd < 0) return fail(errno == EEXIST ? NGPD_IO_ERR_EXISTS : NGPD_IO_ERR_OPEN, "cannot create %s: %s", path, strerror(errno));
    if (threads <= 0) threads = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    threads = (int)std::max<int64_t>(1, std::min<int64_t>(threads, count / 4096 + 1));
    static const char header[] = "# File made by Ruben Band\n";
    int rc = write_all(fd, header, sizeof(header) - 1) ? 0 : fail(NGPD_IO_ERR_OPEN, "write failed: %s", strerror(errno));
    if (!rc) rc = write_records(fd, "v", v, count, threads);
    if (!rc && n) rc = write_records(fd, "vn", n, count, threads);
    if (close(fd) != 0 && !rc) rc = fail(NGPD_IO_ERR_OPEN, "close failed: %s", strerror(errno));
    return rc;
}

NGPD_IO_API int ngpd_io_write_obj(const char* path, const float* v, const float* n, int64_t count, int exclusive, int threads) {
    return write_obj(path, v, n, count, exclusive, threads);
}
NGPD_IO_API int ngpd_io_write_obj_f64(const char* path, const double* v, const double* n, int64_t count, int exclusive, int threads) {
    return write_obj(path, v, n, count, exclusive, threads);
}
