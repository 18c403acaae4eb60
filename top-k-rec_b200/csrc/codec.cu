// NEXT-2 of SURVEY 8(f): the reference's text formats at native speed (host code, multi-threaded).
//   .dat     utils.py:28-55 / evaluate.py:19-28   text matrix, '%f ' per element, '\n' per row
//   ratings  utils.py:58-89 / evaluate.py:30-45   "uid,iid:like,iid:like,..." per line, ids = opaque strings that the
//                                                  id files (one per line) map to row numbers
// The writer is byte-identical to the reference's ('%f' of the value widened to double -- glibc's printf is exact);
// the reader performs the same two roundings as np.float32(token): decimal -> nearest double -> nearest float.
#include "common.cuh"
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

namespace tkr {

struct MappedFile {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); fd = -1; return false; }
        n = (size_t)st.st_size;
        if (n == 0) { p = ""; return true; }
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { ::close(fd); fd = -1; return false; }
        p = (const char*)m;
        return true;
    }
    ~MappedFile() {
        if (p != nullptr && n > 0) munmap((void*)p, n);
        if (fd >= 0) ::close(fd);
    }
};

static int n_threads(size_t work_items) {
    unsigned hc = std::thread::hardware_concurrency();
    size_t t = hc == 0 ? 4 : hc;
    if (t > 32) t = 32;
    if (t > work_items) t = work_items ? work_items : 1;
    return (int)t;
}

static inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }

// Line starts of a text file (a trailing fragment without '\n' counts as a line, as readlines() does).
static void line_offsets(const char* p, size_t n, std::vector<size_t>* off) {
    off->clear();
    size_t pos = 0;
    while (pos < n) {
        off->push_back(pos);
        const void* nl = memchr(p + pos, '\n', n - pos);
        if (nl == nullptr) break;
        pos = (size_t)((const char*)nl - p) + 1;
    }
    off->push_back(n);   // sentinel: line l is [off[l], off[l+1])
}

// One decimal token -> float with the reference's roundings.  Fast path: plain [-]digits[.digits] with < 16
// significant digits is an exact integer over an exact power of ten, and IEEE division rounds it correctly.
static inline float parse_float(const char* b, const char* e) {
    static const double p10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* q = b;
    bool neg = false;
    if (q < e && (*q == '-' || *q == '+')) { neg = *q == '-'; ++q; }
    unsigned long long m = 0;
    int digits = 0, frac = 0;
    bool seen_dot = false, ok = q < e;
    for (; q < e; ++q) {
        const char c = *q;
        if (c >= '0' && c <= '9') {
            if (m != 0 || c != '0') ++digits;
            m = m * 10 + (unsigned)(c - '0');
            if (seen_dot) ++frac;
            if (digits > 15) { ok = false; break; }
        } else if (c == '.' && !seen_dot) {
            seen_dot = true;
        } else { ok = false; break; }
    }
    if (ok && frac <= 22) {
        const double v = (double)m / p10[frac];
        return (float)(neg ? -v : v);
    }
    char buf[64];
    const size_t len = (size_t)(e - b) < sizeof(buf) - 1 ? (size_t)(e - b) : sizeof(buf) - 1;
    memcpy(buf, b, len); buf[len] = 0;
    return (float)strtod(buf, nullptr);
}

}  // namespace tkr

using namespace tkr;

extern "C" int tkr_dat_shape(const char* path, int64_t* rows, int64_t* cols) {
    TKR_CHECK_ARG(path && rows && cols, "NULL argument");
    MappedFile f;
    if (!f.open(path)) { set_error("cannot open %s: %s", path, strerror(errno)); return TKR_ERR_INVALID; }
    std::vector<size_t> off;
    line_offsets(f.p, f.n, &off);
    *rows = (int64_t)off.size() - 1;
    int64_t c = 0;
    if (*rows > 0) {
        const char* q = f.p + off[0]; const char* e = f.p + off[1];
        while (q < e) {
            while (q < e && is_space(*q)) ++q;
            if (q < e) { ++c; while (q < e && !is_space(*q)) ++q; }
        }
    }
    *cols = c;
    return TKR_OK;
}

extern "C" int tkr_dat_read(const char* path, float* out, int64_t rows, int64_t cols) {
    TKR_CHECK_ARG(path && out && rows >= 0 && cols >= 0, "bad arguments");
    MappedFile f;
    if (!f.open(path)) { set_error("cannot open %s: %s", path, strerror(errno)); return TKR_ERR_INVALID; }
    std::vector<size_t> off;
    line_offsets(f.p, f.n, &off);
    if ((int64_t)off.size() - 1 != rows) { set_error("%s has %zu lines, expected %lld", path, off.size() - 1, (long long)rows); return TKR_ERR_INVALID; }
    const int nt = n_threads((size_t)rows);
    std::vector<long long> bad(nt, -1);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        th.emplace_back([&, t]() {
            const int64_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
            for (int64_t r = r0; r < r1; ++r) {
                const char* q = f.p + off[r]; const char* e = f.p + off[r + 1];
                int64_t c = 0;
                while (q < e) {
                    while (q < e && is_space(*q)) ++q;
                    if (q >= e) break;
                    const char* b = q;
                    while (q < e && !is_space(*q)) ++q;
                    if (c < cols) out[r * cols + c] = parse_float(b, q);
                    ++c;
                }
                if (c != cols && bad[t] < 0) bad[t] = r;
            }
        });
    }
    for (auto& x : th) x.join();
    for (int t = 0; t < nt; ++t)
        if (bad[t] >= 0) { set_error("%s: line %lld does not have %lld columns", path, bad[t], (long long)cols); return TKR_ERR_INVALID; }
    return TKR_OK;
}

template <typename T>
static int dat_write_impl(const char* path, const T* mat, int64_t rows, int64_t cols) {
    TKR_CHECK_ARG(path && (mat || rows * cols == 0) && rows >= 0 && cols >= 0, "bad arguments");
    FILE* fp = fopen(path, "wb");
    if (fp == nullptr) { set_error("cannot create %s: %s", path, strerror(errno)); return TKR_ERR_INVALID; }
    const int64_t chunk = 4096;                      // rows formatted per task; tasks are written in order
    const int nt = n_threads((size_t)((rows + chunk - 1) / chunk));
    int rc = TKR_OK;
    for (int64_t base = 0; base < rows && rc == TKR_OK; base += chunk * nt) {
        std::vector<std::string> bufs(nt);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) {
            th.emplace_back([&, t]() {
                const int64_t r0 = base + chunk * t, r1 = r0 + chunk < rows ? r0 + chunk : rows;
                if (r0 >= rows) return;
                std::string& s = bufs[t];
                s.reserve((size_t)(r1 - r0) * (size_t)(cols * 10 + 1));
                char tmp[400];                        // '%f' of FLT_MAX is 47 characters
                for (int64_t r = r0; r < r1; ++r) {
                    for (int64_t c = 0; c < cols; ++c) {
                        const int n = snprintf(tmp, sizeof(tmp), "%f ", (double)mat[r * cols + c]);
                        s.append(tmp, (size_t)n);
                    }
                    s.push_back('\n');
                }
            });
        }
        for (auto& x : th) x.join();
        for (int t = 0; t < nt; ++t)
            if (!bufs[t].empty() && fwrite(bufs[t].data(), 1, bufs[t].size(), fp) != bufs[t].size()) { set_error("write to %s failed: %s", path, strerror(errno)); rc = TKR_ERR_INVALID; break; }
    }
    if (fclose(fp) != 0 && rc == TKR_OK) { set_error("closing %s failed: %s", path, strerror(errno)); rc = TKR_ERR_INVALID; }
    return rc;
}

extern "C" int tkr_dat_write(const char* path, const float* mat, int64_t rows, int64_t cols) { return dat_write_impl(path, mat, rows, cols); }
// float64 input is formatted from the doubles themselves, as the reference's "'%f ' % x" does for a float64 matrix
// (CER keeps E in float64, single/cer.py:27,64,81-85): rounding to float32 first changes the last printed digit of ~2 % of the values.
extern "C" int tkr_dat_write_f64(const char* path, const double* mat, int64_t rows, int64_t cols) { return dat_write_impl(path, mat, rows, cols); }

// Rating file -> flat arrays.  Pass 1 (pairs == NULL): counts only.  line_user[l] = row of the line's user id in
// uid_path (-1 unknown); pair_item[p] = row of the item id in iid_path (-1 unknown); pair_like[p] = 1 iff the label
// text is exactly "1" (the reference compares strings, utils.py:67).  Lines keep file order; a line "uid" alone has
// zero pairs.
extern "C" int tkr_ratings_parse(const char* ratings_path, const char* uid_path, const char* iid_path, int64_t* n_lines,
                                 int64_t* n_pairs, int32_t* line_user, int64_t* line_indptr, int32_t* pair_item,
                                 int8_t* pair_like) {
    TKR_CHECK_ARG(ratings_path && uid_path && iid_path && n_lines && n_pairs, "NULL argument");
    MappedFile fr, fu, fi;
    if (!fr.open(ratings_path)) { set_error("cannot open %s: %s", ratings_path, strerror(errno)); return TKR_ERR_INVALID; }
    std::vector<size_t> off;
    line_offsets(fr.p, fr.n, &off);
    const int64_t L = (int64_t)off.size() - 1;
    auto strip = [](const char*& b, const char*& e) { while (b < e && is_space(*b)) ++b; while (e > b && is_space(e[-1])) --e; };
    if (line_user == nullptr) {                      // sizing pass
        int64_t pairs = 0;
        for (int64_t l = 0; l < L; ++l)
            for (const char* q = fr.p + off[l]; q < fr.p + off[l + 1]; ++q) pairs += *q == ',';
        *n_lines = L; *n_pairs = pairs;
        return TKR_OK;
    }
    TKR_CHECK_ARG(line_indptr && pair_item && pair_like, "NULL output arrays");
    if (*n_lines != L) { set_error("n_lines changed between the sizing and the fill pass"); return TKR_ERR_INVALID; }
    auto load_ids = [&](MappedFile& f, const char* path, std::unordered_map<std::string_view, int32_t>* m) -> bool {
        if (!f.open(path)) { set_error("cannot open %s: %s", path, strerror(errno)); return false; }
        std::vector<size_t> o;
        line_offsets(f.p, f.n, &o);
        m->reserve(o.size() * 2);
        for (size_t l = 0; l + 1 < o.size(); ++l) {
            const char* b = f.p + o[l]; const char* e = f.p + o[l + 1];
            strip(b, e);
            (*m)[std::string_view(b, (size_t)(e - b))] = (int32_t)l;      // later duplicates win, as in the dict comprehension
        }
        return true;
    };
    std::unordered_map<std::string_view, int32_t> um, im;
    if (!load_ids(fu, uid_path, &um) || !load_ids(fi, iid_path, &im)) return TKR_ERR_INVALID;
    // indptr from comma counts, then the lines are independent
    line_indptr[0] = 0;
    for (int64_t l = 0; l < L; ++l) {
        int64_t c = 0;
        for (const char* q = fr.p + off[l]; q < fr.p + off[l + 1]; ++q) c += *q == ',';
        line_indptr[l + 1] = line_indptr[l] + c;
    }
    if (line_indptr[L] != *n_pairs) { set_error("n_pairs changed between the sizing and the fill pass"); return TKR_ERR_INVALID; }
    const int nt = n_threads((size_t)L);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        th.emplace_back([&, t]() {
            for (int64_t l = L * t / nt; l < L * (t + 1) / nt; ++l) {
                const char* b = fr.p + off[l]; const char* e = fr.p + off[l + 1];
                strip(b, e);
                const char* c = (const char*)memchr(b, ',', (size_t)(e - b));
                const char* ue = c ? c : e;
                auto it = um.find(std::string_view(b, (size_t)(ue - b)));
                line_user[l] = it == um.end() ? -1 : it->second;
                int64_t p = line_indptr[l];
                while (c != nullptr) {
                    const char* tb = c + 1;
                    c = (const char*)memchr(tb, ',', (size_t)(e - tb));
                    const char* te = c ? c : e;
                    const char* colon = (const char*)memchr(tb, ':', (size_t)(te - tb));
                    const char* ie = colon ? colon : te;
                    auto jt = im.find(std::string_view(tb, (size_t)(ie - tb)));
                    pair_item[p] = jt == im.end() ? -1 : jt->second;
                    // label = text between the first and the second ':' (split(':')[1])
                    int8_t like = 0;
                    if (colon != nullptr) {
                        const char* lb = colon + 1;
                        const char* c2 = (const char*)memchr(lb, ':', (size_t)(te - lb));
                        const char* le = c2 ? c2 : te;
                        like = (le - lb == 1 && *lb == '1') ? 1 : 0;
                    }
                    pair_like[p] = like;
                    ++p;
                }
            }
        });
    }
    for (auto& x : th) x.join();
    return TKR_OK;
}
