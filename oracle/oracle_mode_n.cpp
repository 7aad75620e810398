// oracle_mode_n.cpp — TEST INFRASTRUCTURE (see oracle_common.h).  North-star stages; filled in below.
#include "oracle_common.h"
extern "C" int orc_voxelize_n(f184o_ctx* c, const f184_view_constants*) { c->err = "unimplemented"; return F184_ERR_UNIMPLEMENTED; }
extern "C" int orc_inject_n(f184o_ctx* c, const f184_sun*, const f184_extended_matrices*) { c->err = "unimplemented"; return F184_ERR_UNIMPLEMENTED; }
extern "C" int orc_mips_n(f184o_ctx* c) { c->err = "unimplemented"; return F184_ERR_UNIMPLEMENTED; }
extern "C" int orc_trace_n(f184o_ctx* c, const f184_trace_constants*) { c->err = "unimplemented"; return F184_ERR_UNIMPLEMENTED; }
