/*
 * The drop-in from plain C: build a model for a four-tip tree from four reference sequences (cls_model_build), upload it
 * (cls_index_create), and run the reference's use-case on a FASTA file (cls_place_sequences) - result records go to
 * <out>.yaml, error texts to <out>.error, exactly as classeq_core::use_cases::place_sequences writes them.
 *
 *   gcc -std=c99 -Iinclude examples/place_fasta.c -Lclasseq2_b200 -l:libclasseq_b200.so -Wl,-rpath,$PWD/classeq2_b200 -o place_fasta
 *   ./place_fasta queries.fasta out          (needs a B200: there is no CPU fallback - without a device it says so and exits 3)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "classeq_b200.h"

#define CHECK(call)                                                          \
    do {                                                                     \
        int rc_ = (call);                                                    \
        if (rc_ != CLS_OK) {                                                 \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, cls_last_error()); \
            return rc_ == CLS_ERR_CUDA ? 3 : 1;                              \
        }                                                                    \
    } while (0)

int main(int argc, char **argv) {
    /* ((t0,t1)n1,(t2,t3)n2)root - nodes in pre-order, node 0 = root */
    static const uint64_t node_id[7] = {0, 1, 2, 3, 4, 5, 6};
    static const uint8_t node_kind[7] = {CLS_KIND_ROOT, CLS_KIND_NODE, CLS_KIND_LEAF, CLS_KIND_LEAF, CLS_KIND_NODE, CLS_KIND_LEAF, CLS_KIND_LEAF};
    static const uint64_t child_off[8] = {0, 2, 4, 4, 4, 6, 6, 6};
    static const uint64_t child_idx[6] = {1, 4, 2, 3, 5, 6};
    static const uint64_t tip_node[4] = {2, 3, 5, 6};
    /* four related reference sequences of 80 bases: the two clades differ in their second half */
    const char *refs[4] = {
        "ACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCCCCGGGTACCGAGCTCGAATTCACTGGCCGTCGTTTTACA",
        "ACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCCCCGGGTACCGAGCTCGAATTCACTGGCCGTCGTTTAACA",
        "ACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCGGTTAACCGGTTAACCGTACGTACGATCGATCGGCTAGCT",
        "ACGTTGCAAGCTTGCATGCCTGCAGGTCGACTCTAGAGGATCGGTTAACCGGTTAACCGTACGTACGATCGATCGGCTAGGT"};
    char bases[4 * 82 + 1] = "";
    uint64_t offsets[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) { strcat(bases, refs[i]); offsets[i + 1] = strlen(bases); }

    cls_model_view tree;
    memset(&tree, 0, sizeof tree);
    tree.k_size = 35; tree.m_size = 4;
    tree.n_nodes = 7; tree.node_id = node_id; tree.node_kind = node_kind; tree.child_off = child_off; tree.child_idx = child_idx;
    cls_built_model *built = NULL;
    CHECK(cls_model_build(&tree, 4, tip_node, (const uint8_t *)bases, offsets, &built));
    cls_model_view model;
    CHECK(cls_built_model_view(built, &tree, &model));
    printf("model: %llu k-mer entries, %llu node sets\n", (unsigned long long)model.n_entries, (unsigned long long)model.n_sets);

    /* the serde fields of the clades, for the record writer */
    static const int64_t parent_id[7] = {-1, 0, 1, 1, 0, 4, 4};
    static const uint8_t children_some[7] = {1, 1, 0, 0, 1, 0, 0}, has_name[7] = {0, 0, 1, 1, 0, 1, 1};
    double support[7], length[7];
    for (int i = 0; i < 7; ++i) { support[i] = node_kind[i] == CLS_KIND_NODE ? 100.0 : NAN; length[i] = i ? 0.01 : 0.0; }
    static const char names[] = "tip_atip_btip_ctip_d";
    static const uint64_t name_off[8] = {0, 0, 0, 5, 10, 10, 15, 20};
    cls_record_tree rt;
    memset(&rt, 0, sizeof rt);
    rt.n_nodes = 7; rt.node_id = node_id; rt.parent_id = parent_id; rt.node_kind = node_kind; rt.children_some = children_some;
    rt.support = support; rt.length = length; rt.has_name = has_name; rt.name_off = name_off; rt.names = names;
    rt.child_off = child_off; rt.child_idx = child_idx;

    if (argc < 3) { fprintf(stderr, "usage: %s queries.fasta out_path [device_mask]\n", argv[0]); return 2; }
    cls_index *index = NULL;
    if (argc > 3) {
        /* one handle over several GPUs: bit d of the mask = CUDA device d ("0" = every visible device); the batches of
         * cls_place_sequences are then cut over them inside the library */
        CHECK(cls_index_create_multi(&model, strtoull(argv[3], NULL, 0), &index));
    } else {
        CHECK(cls_index_create(&model, 0, &index));
    }
    cls_params params;
    cls_params_default(&params);
    uint64_t n = 0;
    CHECK(cls_place_sequences(index, &rt, argv[1], argv[2], &params, 0 /* yaml */, 1 /* overwrite */, &n));
    printf("%llu queries placed\n", (unsigned long long)n);
    cls_index_destroy(index);
    cls_built_model_destroy(built);
    return 0;
}
