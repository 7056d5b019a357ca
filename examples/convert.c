/* examples/convert.c -- the C ABI of include/g2n.h from plain C: GFA file -> CSR arrays + node names.
 *
 *     gcc -I include -o convert examples/convert.c -L gfa2network_b200 -lg2n -Wl,-rpath,$PWD/gfa2network_b200
 *     ./convert graph.gfa
 *
 * The same calls are what a cgo / JNI / N-API binding would make; gfa2network_b200/_capi.py is the ctypes one.
 * Replaces, for one file, parse_gfa(path, build_matrix=True, return_node_list=True) + convert_format(A, "csr")
 * of the reference (builders.py:30-50, utils.py:40-63). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "g2n.h"

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s graph.gfa\n", argv[0]);
        return 2;
    }
    g2n_handle *h = NULL;
    if (g2n_create(0, &h) != G2N_OK) {
        fprintf(stderr, "no usable CUDA device (there is no CPU fallback)\n");
        return 1;
    }
    g2n_params p;
    memset(&p, 0, sizeof p);
    p.directed = 1;               /* parse_gfa defaults: directed, float64 */
    p.dtype = G2N_DTYPE_F64;
    p.want_format = G2N_FMT_CSR;  /* convert_format(A, "csr") fused into the build */
    int rc = g2n_build_file(h, argv[1], &p);
    g2n_diag d;
    g2n_status(h, &d);
    if (d.unknown_byte >= 0) fprintf(stderr, "warning: Skipping unsupported record: %c\n", d.unknown_byte);
    if (rc == G2N_ERR_PARSE) {
        fprintf(stderr, "malformed record (kind %d) in the line at byte %llu\n", d.err_kind, (unsigned long long)d.err_offset);
        g2n_destroy(h);
        return 1;
    }
    if (rc != G2N_OK) {
        fprintf(stderr, "libg2n status %d: %s\n", rc, g2n_last_error(h));
        g2n_destroy(h);
        return 1;
    }
    g2n_sizes_t s;
    g2n_sizes(h, &s);
    int32_t *indptr = malloc((s.n_nodes + 1) * sizeof *indptr);
    int32_t *indices = malloc((s.nnz ? s.nnz : 1) * sizeof *indices);
    double *data = malloc((s.nnz ? s.nnz : 1) * sizeof *data);
    uint64_t names_bytes = 0;
    g2n_names_bytes(h, &names_bytes);
    uint8_t *names = malloc(names_bytes ? names_bytes : 1);
    uint64_t *offsets = malloc((s.n_nodes + 1) * sizeof *offsets);
    if (g2n_fetch_matrix(h, indptr, indices, data) != G2N_OK || g2n_fetch_names(h, names, offsets) != G2N_OK) {
        fprintf(stderr, "fetch failed: %s\n", g2n_last_error(h));
        return 1;
    }
    printf("nodes %llu  nnz %llu  records %llu  device time %.3f ms (%u kernel launches)\n", (unsigned long long)s.n_nodes,
           (unsigned long long)s.nnz, (unsigned long long)d.n_records, d.ms_total, d.gpu_launches);
    for (uint64_t i = 0; i < s.n_nodes && i < 5; i++) {
        printf("  %llu\t%.*s\t->", (unsigned long long)i, (int)(offsets[i + 1] - offsets[i]), (const char *)names + offsets[i]);
        for (int32_t e = indptr[i]; e < indptr[i + 1]; e++) printf(" %d(%g)", indices[e], data[e]);
        printf("\n");
    }
    free(indptr); free(indices); free(data); free(names); free(offsets);
    g2n_destroy(h);
    return 0;
}
