/* A plain-C consumer of the drop-in boundary (include/ultra_rspmm.h), as a cgo / JNI / N-API stub would be: compiled
 * with gcc -std=c99 against the header and linked to libultra_rspmm.so.  Only host-side entry points are called, so it
 * runs without a GPU (tests/test_boundary_cpu.py::test_plain_c_consumer builds and runs it). */
#include <stdio.h>
#include <string.h>

#include "ultra_rspmm.h"

int main(void) {
    const char *version = ultra_rspmm_version();
    if (version == NULL || strncmp(version, "ultra_rspmm sm_100a", 19) != 0) {
        fprintf(stderr, "unexpected version string\n");
        return 1;
    }
    /* FB15k-237 scale: the plan workspace is a pure host-side size computation */
    size_t ws = ultra_rspmm_plan_workspace_bytes(14541, 544230, 474);
    if (ws == 0 || ws > ((size_t)1 << 30)) {
        fprintf(stderr, "implausible workspace size %zu\n", ws);
        return 2;
    }
    if (ultra_rspmm_set_tunable("l2_percent", 25) != 0) return 3;
    if (ultra_rspmm_plan_set_sm_limit(NULL, 140) == 0) return 4; /* null plan is an error code, not a crash */
    /* errors come back as codes plus a message, never as C++ exceptions */
    if (ultra_rspmm_set_tunable("no_such_knob", 1) == 0) return 5;
    if (strstr(ultra_rspmm_last_error(), "unknown tunable") == NULL) return 6;
    printf("%s workspace=%zu\n", version, ws);
    return 0;
}
