// Host-side output writers of the evaluation driver: no kernels here, plain C++ behind the C ABI.
//
//   etch_write_points_vector_ply  <-  utils.GT_utils.save_points_with_vector   src/utils/GT_utils.py:22-55
//       (called twice per scan by src/eval.py:147-149).  The reference writes the ASCII PLY with a Python loop and one
//       f-string per line; the numbers are numpy float32 scalars.  This writer produces the SAME BYTES as that loop does under
//       the numpy in use (two text styles, see fmt_np_float32): std::to_chars gives the same shortest digit strings.
#include "common.cuh"

#include <charconv>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

// Text form of a float32 scalar inside an f-string, `f"{p[0]}"` (GT_utils.py:49-52).  It depends on the numpy the reference runs with:
//   style 0  numpy >= 2:  format(np.float32, "") widens to a Python float -> repr(float(x)): shortest digits that round-trip the DOUBLE,
//            scientific iff the decimal exponent is < -4 or >= 16                      ("99999.8984375", "9.999999747378752e-05")
//   style 1  numpy 1.x:   str(np.float32): shortest digits that round-trip the FLOAT32 (dragon4, unique), scientific iff
//            |x| < 1e-4 or |x| >= 1e6                                                   ("99999.9", "1e-04", "1.234567e+06")
// Both print at least one fractional digit in positional form and an exponent of at least two digits.
char* layout(char* out, const char* sci, bool force_sci_rule_by_value, double absval) {
    const char* p = sci;
    if (*p == '-') { *out++ = '-'; ++p; }
    char digits[24];
    int nd = 0;
    for (; *p && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    const int e10 = atoi(p + 1);
    const bool scientific = force_sci_rule_by_value ? (absval < 1e-4 || absval >= 1e6) : (e10 < -4 || e10 >= 16);
    if (scientific) {
        *out++ = digits[0];
        if (nd > 1) { *out++ = '.'; memcpy(out, digits + 1, nd - 1); out += nd - 1; }
        *out++ = 'e';
        *out++ = e10 < 0 ? '-' : '+';
        const int ae = e10 < 0 ? -e10 : e10;
        if (ae < 10) *out++ = '0';
        return std::to_chars(out, out + 4, ae).ptr;
    }
    if (e10 < 0) {                        // 0.000ddd
        *out++ = '0'; *out++ = '.';
        for (int i = 0; i < -e10 - 1; ++i) *out++ = '0';
        memcpy(out, digits, nd);
        return out + nd;
    }
    const int ni = e10 + 1;               // digits in front of the point
    for (int i = 0; i < ni; ++i) *out++ = i < nd ? digits[i] : '0';
    *out++ = '.';
    if (nd > ni) { memcpy(out, digits + ni, nd - ni); out += nd - ni; }
    else *out++ = '0';
    return out;
}

char* fmt_np_float32(char* out, float x, int style) {
    if (x != x) { memcpy(out, "nan", 3); return out + 3; }
    if (x == INFINITY) { memcpy(out, "inf", 3); return out + 3; }
    if (x == -INFINITY) { memcpy(out, "-inf", 4); return out + 4; }
    if (x == 0.0f) {
        if (std::signbit(x)) *out++ = '-';
        memcpy(out, "0.0", 3);
        return out + 3;
    }
    char sci[40];
    if (style == 1) {
        *std::to_chars(sci, sci + sizeof(sci), x, std::chars_format::scientific).ptr = 0;           // shortest float32 round trip
        return layout(out, sci, true, fabs((double)x));
    }
    *std::to_chars(sci, sci + sizeof(sci), (double)x, std::chars_format::scientific).ptr = 0;       // shortest double round trip
    return layout(out, sci, false, 0.0);
}

}  // namespace

// self-test hook for the CPU suite: formats n floats, '\n'-separated, into out (capacity cap); returns bytes written or -1
ETCH_API long long etch_format_np_float32(const float* x, int n, int style, char* out, long long cap) {
    char* o = out;
    for (int i = 0; i < n; ++i) {
        if (o - out + 40 > cap) return -1;
        o = fmt_np_float32(o, x[i], style);
        *o++ = '\n';
    }
    return (long long)(o - out);
}

// HOST pointers: hit_points [n,3], vectors [n,3] float32 (row-major).  Returns 0, or ETCH_EINVAL / the errno of a failed write.
ETCH_API int etch_write_points_vector_ply(const char* path, const float* hit_points, const float* vectors, int n, int style) {
    if (!path || !hit_points || !vectors || n < 0 || style < 0 || style > 1) return ETCH_EINVAL;
    std::string buf;
    buf.reserve((size_t)n * 2 * 48 + (size_t)n * 16 + 512);
    buf += "ply\nformat ascii 1.0\n";
    buf += "element vertex " + std::to_string(2 * (long long)n) + "\n";
    buf += "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n";
    buf += "element edge " + std::to_string(n) + "\n";
    buf += "property int vertex1\nproperty int vertex2\nend_header\n";
    char line[160];
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < n; ++i) {
            char* o = line;
            for (int c = 0; c < 3; ++c) {
                const float h = hit_points[(size_t)i * 3 + c];
                const float v = pass == 0 ? h : h - vectors[(size_t)i * 3 + c];     // vector_end_points = hit_points - vectors (float32)
                o = fmt_np_float32(o, v, style);
                *o++ = ' ';
            }
            const char* tail = pass == 0 ? "255 0 0\n" : "0 0 255\n";
            const size_t tl = strlen(tail);
            memcpy(o, tail, tl);
            buf.append(line, (size_t)(o - line) + tl);
        }
    for (int i = 0; i < n; ++i) {
        char* o = std::to_chars(line, line + 16, i).ptr;
        *o++ = ' ';
        o = std::to_chars(o, o + 16, n + i).ptr;
        *o++ = '\n';
        buf.append(line, (size_t)(o - line));
    }
    FILE* f = fopen(path, "wb");
    if (!f) return errno ? errno : ETCH_EINVAL;
    const size_t w = fwrite(buf.data(), 1, buf.size(), f);
    const int rc = fclose(f);
    return (w == buf.size() && rc == 0) ? ETCH_OK : (errno ? errno : ETCH_EINVAL);
}
