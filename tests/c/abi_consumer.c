/* A plain C99 consumer of the drop-in boundary: include/riemann_b200.h must be valid C (no C++ in the signatures),
 * the library must link from C, and bad arguments must come back as RMN_ERR_PARAM with a message -- none of which
 * needs a GPU.  Built and run by tests/test_abi_and_host.py. */
#include "riemann_b200.h"

#include <stdio.h>
#include <string.h>

int main(void) {
    rmn_model_t* m = NULL;
    double mu[2] = {0.0, 0.0};
    double prec[4] = {1.0, 0.0, 0.0, 1.0};
    int rc;
    if (rmn_version() != 100) { printf("version %d\n", rmn_version()); return 1; }
    rc = rmn_model_gaussian_create(&m, 0, mu, prec, prec, 0.0);                 /* d = 0 is a bad argument */
    if (rc != RMN_ERR_PARAM || strlen(rmn_last_error()) == 0) { printf("rc %d\n", rc); return 2; }
    rc = rmn_model_gaussian_create(NULL, 2, mu, prec, prec, 0.0);               /* no place for the handle */
    if (rc != RMN_ERR_PARAM) { printf("rc %d\n", rc); return 3; }
    if (rmn_sampler_destroy(NULL) != RMN_OK && strlen(rmn_last_error()) == 0) return 4;
    printf("abi consumer ok: %s\n", rmn_last_error());
    return 0;
}
