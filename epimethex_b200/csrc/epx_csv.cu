/*
 * epx_csv.cu — host-only: the reference's result files, byte-compatible with R's write.csv
 * (R/EpimethexAnalysis.R:470 CG_data_individually.csv, :545 CG_of_a_genes.csv, :577 CG_by_position.csv,
 * :615 CG_by_islands.csv), formatted on several threads.  SURVEY 8f rank 3: at full scale the table has
 * 474,611 rows x 18 columns and an interpreted writer dominates the end-to-end time.
 *
 * Number rendering follows R's formatReal with 15 significant digits on ONE value at a time (what write.table
 * and as.character do): the fewest significant digits (<= 15) that reproduce the value to 15 digits, fixed
 * notation unless scientific is strictly narrower ("1e-04", "1e+05", but "0.00012345" and "123456").
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/epx.h"

namespace {

/* as R prints a double with digits = 15 */
void format_r15(double x, std::string &out)
{
    if (std::isnan(x)) { out += "NaN"; return; }
    if (std::isinf(x)) { out += x > 0 ? "Inf" : "-Inf"; return; }
    if (x == 0.0) { out += "0"; return; }
    char buf[64];
    snprintf(buf, sizeof buf, "%.14e", x);             /* d.dddddddddddddde[+-]XX */
    const char *e = strchr(buf, 'e');
    const int kp = atoi(e + 1);
    const int neg = buf[0] == '-';
    const char *mant = buf + neg;                      /* "d.dddd..." */
    int nsig = 15;
    for (const char *p = e - 1; p > mant + 1 && *p == '0'; --p) nsig--;   /* strip trailing zeros of the 14 decimals */
    const int left = kp >= 0 ? kp + 1 : 1;
    int rgt = nsig - kp - 1;
    if (rgt < 0) rgt = 0;
    const int wF = neg + left + (rgt ? rgt + 1 : 0);
    const int mF = nsig - 1;
    const int wE = neg + (mF > 0 ? mF + 1 : 0) + 1 + 4 + ((kp >= 100 || kp <= -100) ? 1 : 0);
    /* assemble from the 15 digits already in hand (a second snprintf would double the cost) */
    char dig[16];
    dig[0] = mant[0];
    for (int i = 1; i < 15; i++) dig[i] = mant[1 + i];
    if (neg) out += '-';
    if (wF <= wE) {
        if (kp >= 15) { snprintf(buf, sizeof buf, "%.0f", std::fabs(x)); out += buf; return; } /* digits beyond the 15th */
        if (kp >= 0) {
            for (int i = 0; i <= kp; i++) out += i < nsig ? dig[i] : '0';
            if (rgt) { out += '.'; out.append(dig + kp + 1, (size_t)rgt); }
        } else {
            out += "0.";
            out.append((size_t)(-kp - 1), '0');
            out.append(dig, (size_t)nsig);
        }
    } else {
        out += dig[0];
        if (mF > 0) { out += '.'; out.append(dig + 1, (size_t)mF); }
        out += 'e';
        out += kp < 0 ? '-' : '+';
        const int ak = kp < 0 ? -kp : kp;
        if (ak >= 100) out += (char)('0' + ak / 100);
        out += (char)('0' + (ak / 10) % 10);
        out += (char)('0' + ak % 10);
    }
}

void quoted(const char *s, std::string &out)
{
    out += '"';
    for (; s && *s; ++s) {
        if (*s == '"') out += '"';                     /* write.csv: qmethod = "double" */
        out += *s;
    }
    out += '"';
}

int fail(char *err, size_t errlen, int code, const char *msg)
{
    if (err && errlen) snprintf(err, errlen, "%s", msg);
    return code;
}

} // namespace

extern "C" {

int epx_format_real15(double x, char *buf, size_t buflen)
{
    std::string s;
    format_r15(x, s);
    if (!buf || buflen < s.size() + 1) return EPX_EINVAL;
    memcpy(buf, s.c_str(), s.size() + 1);
    return EPX_OK;
}

int epx_write_csv(const char *path, int64_t n_rows, int n_cols, const char *const *row_names,
                  const char *const *col_names, const double *values, const uint8_t *na,
                  const char *const *text_col, const char *text_col_name, int quote_numbers, int threads,
                  char *err, size_t errlen)
{
    if (!path || n_rows < 0 || n_cols < 0 || !col_names || (n_rows && (!row_names || (n_cols && !values))))
        return fail(err, errlen, EPX_EINVAL, "epx_write_csv: bad arguments");
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if ((int64_t)threads > n_rows / 1024 + 1) threads = (int)(n_rows / 1024 + 1);
    std::vector<std::string> part((size_t)threads);
    auto work = [&](int t) {
        const int64_t r0 = n_rows * t / threads, r1 = n_rows * (t + 1) / threads;
        std::string &o = part[(size_t)t];
        o.reserve((size_t)(r1 - r0) * (size_t)(24 + 20 * n_cols));
        for (int64_t r = r0; r < r1; r++) {
            quoted(row_names[r], o);
            for (int c = 0; c < n_cols; c++) {
                o += ',';
                const double v = values[r * n_cols + c];
                const bool is_na = na ? na[r * n_cols + c] != 0 : std::isnan(v);
                if (is_na) { o += "NA"; continue; }      /* write.csv never quotes NA */
                if (quote_numbers) o += '"';
                format_r15(v, o);
                if (quote_numbers) o += '"';
            }
            if (text_col) {
                o += ',';
                if (text_col[r]) quoted(text_col[r], o); else o += "NA";
            }
            o += '\n';
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();

    FILE *f = fopen(path, "wb");
    if (!f) return fail(err, errlen, EPX_EINVAL, "epx_write_csv: cannot open the output file");
    std::string head = "\"\"";
    for (int c = 0; c < n_cols; c++) { head += ','; quoted(col_names[c], head); }
    if (text_col) { head += ','; quoted(text_col_name ? text_col_name : "", head); }
    head += '\n';
    bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
    for (int t = 0; t < threads && ok; t++) ok = fwrite(part[(size_t)t].data(), 1, part[(size_t)t].size(), f) == part[(size_t)t].size();
    ok = (fclose(f) == 0) && ok;
    return ok ? EPX_OK : fail(err, errlen, EPX_EINVAL, "epx_write_csv: write failed");
}

} // extern "C"
