// TEST INFRASTRUCTURE (oracle harness).  Runs the UNMODIFIED reference prover + verifier (compiled from
// /root/reference by oracle/Makefile) with a seeded challenge stream and records the transcript.
// Mirrors the reference mains (src/main_demo_lenet.cpp:19-40, src/main_demo_vgg.cpp:20-42).
//
//   ref_run lenet  <input.csv> <config.csv> <pic_cnt> <seed> [--transcript out.bin] [--shapes] [--circuit-hash]
//                  [--gens real|degenerate] [--dump-dir DIR]
//   ref_run vgg    <input.csv> <config.csv> <network.csv> <pic_cnt> <seed> [...]
// the circuit builder was compiled against the moved class names (oracle/Makefile RENAME)
#define prover ref_prover
#define polyProver ref_polyProver
#include <neuralNetwork.hpp>
#include <models.hpp>
#undef prover
#undef polyProver
#include <verifier.hpp>
#include "seeded_rng.hpp"
#include <chrono>

vector<std::string> output_tb(16, "");

static uint64_t fnv(const void *p, size_t n, uint64_t h = 0xcbf29ce484222325ULL) {
    auto *b = static_cast<const uint8_t *>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ULL; }
    return h;
}

static void printShapes(const layeredCircuit &C, const vector<vector<F>> &val, bool hashes) {
    static const char *names[] = {"INPUT", "FFT", "IFFT", "ADD_BIAS", "RELU", "Sqr", "OPT_AVG_POOL", "MAX_POOL", "AVG_POOL",
                                  "DOT_PROD", "PADDING", "FCONN", "NCONV", "NCONV_MUL", "NCONV_ADD"};
    for (int i = 0; i < C.size; ++i) {
        auto &c = C.circuit[i];
        printf("L %d %s size %u bl %d u0 %u %d u1 %u %d v0 %u %d v1 %u %d mbu %d mbv %d ph2 %d fftbl %d zsi %u uni %zu bin %zu",
               i, names[(int) c.ty], c.size, (int) c.bit_length, c.size_u[0], (int) c.bit_length_u[0], c.size_u[1],
               (int) c.bit_length_u[1], c.size_v[0], (int) c.bit_length_v[0], c.size_v[1], (int) c.bit_length_v[1],
               (int) c.max_bl_u, (int) c.max_bl_v, (int) c.need_phase2, (int) c.fft_bit_length, c.zero_start_id,
               c.uni_gates.size(), c.bin_gates.size());
        if (hashes) {
            // gate arrays are hashed field by field so that struct padding does not leak in
            uint64_t hu = 0xcbf29ce484222325ULL, hb = hu;
            for (auto &g : c.uni_gates) { u32 t[4] = {g.g, g.u, g.lu, g.sc}; hu = fnv(t, sizeof t, hu); }
            for (auto &g : c.bin_gates) { u32 t[5] = {g.g, g.u, g.v, g.sc, g.l}; hb = fnv(t, sizeof t, hb); }
            uint64_t hou = fnv(c.ori_id_u.data(), c.ori_id_u.size() * 4), hov = fnv(c.ori_id_v.data(), c.ori_id_v.size() * 4);
            uint64_t hv = 0xcbf29ce484222325ULL;
            for (auto &x : val[i]) { uint8_t b[32]; x.serialize(b, 32); hv = fnv(b, 32, hv); }
            uint8_t sb[32]; c.scale.serialize(sb, 32);
            printf(" h_uni %016lx h_bin %016lx h_oriu %016lx h_oriv %016lx h_val %016lx nval %zu h_scale %016lx",
                   hu, hb, hou, hov, hv, val[i].size(), fnv(sb, 32));
        }
        printf("\n");
    }
}

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: see header\n"); return 2; }
    initPairing(mcl::BLS12_381);
    std::string model = argv[1];
    int k = 2;
    std::string in_file = argv[k++], conf_file = argv[k++], net_file;
    if (model == "vgg") net_file = argv[k++];
    int pic_cnt = atoi(argv[k++]);
    uint64_t seed = strtoull(argv[k++], nullptr, 0);
    bool real_gens = false;
    std::string tr_out; bool shapes = false, hashes = false;
    for (; k < argc; ++k) {
        std::string a = argv[k];
        if (a == "--transcript") tr_out = argv[++k];
        else if (a == "--gens") real_gens = std::string(argv[++k]) == "real";
        else if (a == "--shapes") shapes = true;
        else if (a == "--circuit-hash") shapes = hashes = true;
        else if (a == "--dump-dir" && k + 1 < argc) {   // per-call table hashes (record_proxy.hpp: dumpTables) -> DIR/tables.txt
            std::string path = std::string(argv[++k]) + "/tables.txt";
            table_dump_file() = fopen(path.c_str(), "w");
            if (!table_dump_file()) { fprintf(stderr, "cannot write %s\n", path.c_str()); return 3; }
        }
    }

    if (real_gens) install_real_base_point();
    SeededStream rng(seed);
    rng.install();

    auto t0 = std::chrono::steady_clock::now();
    ref_prover rp;
    std::unique_ptr<neuralNetwork> nn;
    if (model == "lenet") nn.reset(new lenet(32, 32, 1, pic_cnt, MAX, in_file, conf_file, ""));
    else if (model == "vgg") nn.reset(new vgg(32, 32, 3, pic_cnt, in_file, conf_file, "", net_file));
    else { fprintf(stderr, "unknown model\n"); return 2; }
    nn->create(rp, false);
    auto t1 = std::chrono::steady_clock::now();
    if (shapes) printShapes(rp.C, rp.val, hashes);

    prover proxy(rp);
    verifier v(&proxy, rp.C);
    bool ok = v.verify();
    auto t2 = std::chrono::steady_clock::now();

    printf("RESULT ok %d n_fr %lu n_g1 %lu bytes %zu fnv %016lx challenges %lu create_s %.3f verify_wall_s %.3f\n", (int) ok,
           transcript().n_fr, transcript().n_g1, transcript().bytes.size(), transcript().fnv1a(), rng.calls,
           std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
    printf("TABLE ");
    for (auto &s : output_tb) printf("%s, ", s.c_str());
    puts("");
    if (table_dump_file()) fclose(table_dump_file());
    if (!tr_out.empty() && !transcript().save(tr_out)) { fprintf(stderr, "cannot write %s\n", tr_out.c_str()); return 3; }
    return ok ? 0 : 1;
}
