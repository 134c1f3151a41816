/* TEST INFRASTRUCTURE ONLY -- minimal stand-in for <htslib/hts.h>.
 * The reference (warp9seq/minimod v0.5.0) includes this header from
 * src/minimod.h:39 but uses nothing from it beyond what sam.h declares. */
#ifndef ORACLE_SHIM_HTS_H
#define ORACLE_SHIM_HTS_H
#include "sam.h"
#endif
