// temporary: north-star stages are added in mode_n_*.cu
#include "f184_internal.h"
int f184_voxelize_n(f184_ctx* c, const f184_view_constants*) { return f184_fail(c, F184_ERR_UNIMPLEMENTED, "mode N voxelize not built yet"); }
int f184_inject_n(f184_ctx* c, const f184_sun*, const f184_extended_matrices*) { return f184_fail(c, F184_ERR_UNIMPLEMENTED, "mode N inject not built yet"); }
int f184_mips_n(f184_ctx* c) { return f184_fail(c, F184_ERR_UNIMPLEMENTED, "mode N mips not built yet"); }
int f184_trace_n(f184_ctx* c, const f184_trace_constants*) { return f184_fail(c, F184_ERR_UNIMPLEMENTED, "mode N trace not built yet"); }
int f184_mode_n_release(f184_ctx*) { return 0; }
