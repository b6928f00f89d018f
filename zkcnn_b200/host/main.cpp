// zkcnn_prove: command-line front end of the stand-alone build; the counterpart of the reference's demo mains
// (src/main_demo_lenet.cpp:19-40, src/main_demo_vgg.cpp:20-42) with a seeded challenge stream and a proof dump.
//
//   zkcnn_prove lenet <input.csv> <pic_cnt> <seed> [options]
//   zkcnn_prove vgg   <input.csv> "<network description>" <pic_cnt> <seed> [options]
// options: --gens real|degenerate   --prover-only (skip the verifier's wiring predicates and G1 checks; full verification is the default)   --csprng (challenges from the OS CSPRNG, like the reference)   --fiat-shamir   --round-by-round (one device call per sumcheck round)   --transcript out.bin   --circuit-hash out.txt
//          --repeat N (prove N times, witness resident after the first)   --device D
#include "../../include/zkcnn_host.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

int main(int argc, char **argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: zkcnn_prove lenet <input> <pic_cnt> <seed> [...] | vgg <input> \"<network>\" <pic_cnt> <seed> [...]\n");
        return 2;
    }
    int k = 1;
    const std::string model = argv[k++];
    const std::string input = argv[k++];
    std::string network;
    if (model == "vgg") network = argv[k++];
    if (k + 2 > argc) { fprintf(stderr, "missing pic_cnt / seed\n"); return 2; }
    const int pic_cnt = atoi(argv[k++]);
    const uint64_t seed = strtoull(argv[k++], nullptr, 0);
    uint32_t flags = 0;
    int repeat = 1, device = 0;
    std::string tr_out, hash_out;
    for (; k < argc; ++k) {
        const std::string a = argv[k];
        if (a == "--gens" && k + 1 < argc) { if (std::string(argv[++k]) == "real") flags |= ZKH_REAL_GENERATORS; }
        else if (a == "--check") flags |= ZKH_CHECK_PREDICATES;   // (the default; kept for old command lines)
        else if (a == "--prover-only") flags |= ZKH_PROVER_ONLY;
        else if (a == "--csprng") flags |= ZKH_CSPRNG_CHALLENGES;
        else if (a == "--fiat-shamir") flags |= ZKH_FIAT_SHAMIR;
        else if (a == "--round-by-round") flags |= ZKH_ROUND_BY_ROUND;
        else if (a == "--transcript" && k + 1 < argc) tr_out = argv[++k];
        else if (a == "--circuit-hash" && k + 1 < argc) hash_out = argv[++k];
        else if (a == "--repeat" && k + 1 < argc) repeat = atoi(argv[++k]);
        else if (a == "--device" && k + 1 < argc) device = atoi(argv[++k]);
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    zkh_session *s = zkh_create(model.c_str(), network.c_str(), pic_cnt, device);
    if (!s) { fprintf(stderr, "zkh_create: %s\n", zkh_last_error()); return 3; }
    if (zkh_input_file(s, input.c_str()) || zkh_build(s)) { fprintf(stderr, "build: %s\n", zkh_last_error()); return 3; }
    if (!hash_out.empty() && zkh_circuit_dump(s, hash_out.c_str(), 1)) { fprintf(stderr, "dump: %s\n", zkh_last_error()); return 3; }
    zkh_stats st;
    int rc = 0;
    for (int i = 0; i < repeat; ++i) {
        if (zkh_prove(s, seed, flags | (i ? ZKH_WITNESS_RESIDENT : 0), &st)) { fprintf(stderr, "prove: %s\n", zkh_last_error()); return 4; }
        printf("RESULT ok %d n_fr %lu n_g1 %lu bytes %lu fnv %016lx challenges %lu prove_s %.4f poly_s %.4f upload_s %.4f wall_s %.4f "
               "verifier_s %.4f launches %lu class %d\n",
               st.ok, (unsigned long) st.n_fr, (unsigned long) st.n_g1, (unsigned long) st.proof_bytes, (unsigned long) st.fnv1a,
               (unsigned long) st.challenges, st.prove_s, st.poly_s, st.upload_s, st.wall_s, st.verifier_s, (unsigned long) st.gpu_launches,
               zkh_inferred_class(s, 0));
        if (!st.ok) rc = 1;
    }
    if (!tr_out.empty()) {
        uint64_t n = 0;
        const uint8_t *b = zkh_proof(s, &n);
        FILE *f = fopen(tr_out.c_str(), "wb");
        if (!f || fwrite(b, 1, n, f) != n) { fprintf(stderr, "cannot write %s\n", tr_out.c_str()); return 5; }
        fclose(f);
    }
    zkh_destroy(s);
    return rc;
}
