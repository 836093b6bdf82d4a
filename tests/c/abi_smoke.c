/* Plain-C caller of the boundary: proves include/ft8_b200.h is valid C99 and that the shared library links from C.
 * Built and run by tests/test_abi.py::test_header_compiles_and_links_from_c (no GPU needed: without a device
 * ft8_create must fail with FT8_E_NODEVICE and a message, never fall back to a CPU path). */
#include <stdio.h>
#include <string.h>
#include "ft8_b200.h"

int main(void) {
    ft8_cfg cfg;
    ft8_handle* h = NULL;
    ft8_record r;
    int rc;
    ft8_default_cfg(&cfg);
    if (cfg.max_cands != 200 || cfg.osd_singleflips != 30 || cfg.osd_doubleflips != 2) return 2;
    if (sizeof(r) != 64 || sizeof(cfg) != 48) return 3;
    cfg.max_cycles = 1;
    rc = ft8_create(0, &cfg, &h);
    if (rc == FT8_OK) {                 /* a GPU is present: exercise one call and tear down */
        uint32_t w[3] = {0u, 0u, 0u};
        int32_t flags = -1;
        void* pinned = NULL;
        if (ft8_host_alloc(4096, FT8_HOST_WRITE_COMBINED, &pinned) != FT8_OK || pinned == NULL) return 6;
        memset(pinned, 0, 4096);
        if (ft8_host_free(pinned) != FT8_OK) return 7;
        rc = ft8_crc14(h, w, 1, &flags, FT8_MEM_HOST);
        printf("gpu present: ft8_crc14 rc=%d flags=%d\n", rc, (int)flags);
        ft8_destroy(h);
        return (rc == FT8_OK && flags == 0) ? 0 : 4;
    }
    {                                   /* no device: the allocator fails cleanly too */
        void* pinned = (void*)1;
        if (ft8_host_alloc(4096, 0, &pinned) == FT8_OK || pinned != NULL) return 8;
        if (ft8_host_free(NULL) != FT8_OK) return 9;
    }
    printf("no device: rc=%d msg=%s\n", rc, ft8_last_error(NULL));
    return (rc == FT8_E_NODEVICE && strlen(ft8_last_error(NULL)) > 0) ? 0 : 5;
}
