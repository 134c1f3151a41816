#include "modopts.h"
#include <ctype.h>
#include <errno.h>
#include <stdlib.h>
#include <string.h>

namespace mmh {

static bool valid_base(char c) { return c && strchr("ACGTUNacgtun", c) != nullptr; }

// default_context[] + get_default_context(), src/mod.c:99-112
const char *default_context(const std::string &code) {
    if (code.size() == 1) {
        switch (code[0]) {
        case '*': return "*";
        case 'm': case 'h': return "CG";
        case 'f': case 'c': case 'C': return "C";
        case 'g': case 'e': case 'b': case 'T': case 'U': return "T";
        case 'a': case 'A': return "A";
        case 'o': case 'G': return "G";
        case 'n': case 'N': return "N";
        default: break;
        }
    }
    return "CG";
}

bool is_tested_case(const std::string &code, const std::string &context) {      // tested_cases[], src/mod.c:101
    static const char *t[] = {"m[CG]", "h[CG]", "m[C]", "h[C]", "m[*]", "h[*]", "*[*]", "21839[C]", "a[A]", "a[*]", "19229[G]",
                              "19229[*]", "69426[A]", "17596[A]", "19228[C]", "19227[T]", "17802[T]", "17802[*]", "e[T]", "b[T]", "m[CT]"};
    std::string s = code + "[" + context + "]";
    for (const char *x : t) if (s == x) return true;
    return false;
}

static std::string fmt(const char *f, const std::string &a) { char b[512]; snprintf(b, sizeof(b), f, a.c_str()); return b; }
static std::string fmtc(const char *f, char c) { char b[512]; snprintf(b, sizeof(b), f, c); return b; }
static std::string fmtcs(const char *f, char c, const std::string &a) { char b[512]; snprintf(b, sizeof(b), f, c, a.c_str()); return b; }

bool parse_mod_codes(const std::string &str, std::vector<ModSpec> *out, std::string *err) {
    const char *s = str.c_str();
    size_t i = 0;
    out->clear();
    while (s[i] != '\0') {
        bool has_nums = false, has_alpha = false;
        ModSpec m;
        while (s[i] != ',' && s[i] != '[' && s[i] != '\0') {
            char c = s[i];
            if (isalpha((unsigned char)c) && (unsigned char)c < 128) has_alpha = true;
            else if (c == '*') has_alpha = true;
            else if (c >= '0' && c <= '9') has_nums = true;
            else { *err = fmtc("Invalid character %c in modification code in -c argument", c); return false; }
            m.code.push_back(c); ++i;
        }
        if (has_alpha && has_nums) { *err = fmt("Modification code %s cannot contain both letters and numbers in -c argument", m.code); return false; }
        if (s[i] == '[') {
            ++i;
            bool is_star = false;
            size_t j = 0;
            while (s[i] != ']') {
                if (s[i] == '*') is_star = true;
                else if (!valid_base(s[i])) {
                    if (s[i] == '\0') *err = fmt("Context not closed with a ] for modification code %s in -c argument", m.code);
                    else *err = fmtcs("Invalid character %c in context for modification code %s in -c argument", s[i], m.code);
                    return false;
                }
                char c = (char)toupper((unsigned char)s[i]);
                if (c == 'U') c = 'T';
                m.context.push_back(c); ++i; ++j;
            }
            if (is_star && j > 1) { *err = fmt("Invalid context for modification code %s. * should be the only character within [ and ] in -c argument", m.code); return false; }
            m.context_given = true;
            ++i;
            if (s[i] == ',') ++i;
        } else if (s[i] == ',') { m.context = default_context(m.code); ++i; }
        else if (s[i] == '\0') { m.context = default_context(m.code); }
        else { *err = fmtcs("Invalid character %c after modification code %s in -c argument", s[i], m.code); return false; }
        for (const ModSpec &o : *out)
            if (o.code == m.code) { *err = fmt("Duplicate modification code %s found in -c argument", m.code); return false; }
        out->push_back(m);
    }
    return true;
}

bool parse_mod_threshes(const std::string &str, std::vector<ModSpec> *mods, std::string *err) {
    const char *s = str.c_str();
    size_t i = 0, n = 0;
    double d = 0.0;
    while (s[i] != '\0') {
        std::string tok;
        while (s[i] != ',' && s[i] != '\0') tok.push_back(s[i++]);
        errno = 0;
        d = atof(tok.c_str());
        if (errno != 0) { *err = fmt("Invalid threshold. You entered %s", tok); return false; }
        if (d < 0 || d > 1) { char b[256]; snprintf(b, sizeof(b), "Modification threshold should be in the range 0.0 to 1.0. You entered %f", d); *err = b; return false; }
        if (n < mods->size()) (*mods)[n].thresh = d;
        ++n;
        if (s[i] == '\0') break;
        ++i;
    }
    if (n == 1) { for (ModSpec &m : *mods) m.thresh = d; }
    else if (n != mods->size()) {
        char b[256];
        snprintf(b, sizeof(b), "Number of modification codes and thresholds do not match. Codes:%d, Thresholds:%d", (int)mods->size(), (int)n);
        *err = b; return false;
    }
    return true;
}

// The integer ML-vs-threshold test, evaluated once per ML byte value with exactly the reference's
// double expressions (THRESH_UINT8_TO_DBL, src/mod.c:56; comparisons src/mod.c:1181-1191).
void build_call_lut(double thresh, uint8_t lut[256]) {
    for (int p = 0; p < 256; ++p) {
        volatile double x = (double)((p + 0.5) / 256.0);
        volatile double one_minus = 1 - thresh;
        if (x >= thresh) lut[p] = MMC_LUT_CALLED | MMC_LUT_MOD;
        else if (x <= one_minus) lut[p] = MMC_LUT_CALLED;
        else lut[p] = 0;
    }
}

bool to_mmc_mods(const std::vector<ModSpec> &mods, std::vector<mmc_mod_t> *out, std::string *err) {
    out->clear();
    if (mods.empty() || mods.size() > MMC_MAX_MODS) { *err = "between 1 and 64 modification codes are supported"; return false; }
    for (const ModSpec &m : mods) {
        mmc_mod_t x;
        memset(&x, 0, sizeof(x));
        if (m.code.empty() || m.code.size() > MMC_MAX_CODE_LEN) { *err = fmt("Modification code '%s' is empty or longer than 8 characters", m.code); return false; }
        if (m.context.empty() || m.context.size() > MMC_MAX_CONTEXT) { *err = fmt("Context '%s' is empty or longer than 32 characters", m.context); return false; }
        memcpy(x.code, m.code.data(), m.code.size());
        memcpy(x.context, m.context.data(), m.context.size());
        build_call_lut(m.thresh, x.call_lut);
        out->push_back(x);
    }
    return true;
}

}  // namespace mmh
