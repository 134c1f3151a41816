// modopts.h -- "-c code[context],..." and "-m thresh,..." parsing with the reference's grammar,
// defaults and error text (parse_mod_codes / parse_mod_threshes, src/mod.c:204-398), producing
// the flat mmc_mod_t table (with the 256-entry call LUT) that the device consumes.
#ifndef MMH_MODOPTS_H
#define MMH_MODOPTS_H
#include <string>
#include <vector>
#include "minimod_cuda.h"
namespace mmh {
struct ModSpec { std::string code, context; double thresh = 0.8; bool context_given = false; };
// returns false and sets *err (the reference's ERROR text) on malformed input
bool parse_mod_codes(const std::string &codes, std::vector<ModSpec> *out, std::string *err);
bool parse_mod_threshes(const std::string &threshes, std::vector<ModSpec> *mods, std::string *err);
void build_call_lut(double thresh, uint8_t lut[256]);
bool to_mmc_mods(const std::vector<ModSpec> &mods, std::vector<mmc_mod_t> *out, std::string *err);
const char *default_context(const std::string &code);
bool is_tested_case(const std::string &code, const std::string &context);
}
#endif
