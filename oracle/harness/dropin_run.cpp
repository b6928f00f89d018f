// TEST INFRASTRUCTURE (oracle harness).  The drop-in check: the reference's OWN verifier, circuit builder and model
// zoo (compiled from /root/reference by oracle/Makefile with -include zkcnn_b200/host/dropin.hpp) drive the
// zkcnn_b200 prover through the reference's public prover interface.  Same command line as ref_run; the transcript it
// records must be byte-identical to ref_run's for the same seed.
#include <models.hpp>
#include <verifier.hpp>
#include <neuralNetwork.hpp>
#include "seeded_rng.hpp"
#include <chrono>

vector<std::string> output_tb(16, "");

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: dropin_run lenet|vgg <input> <config> [<network>] <pic_cnt> <seed> [--transcript out.bin]\n"); return 2; }
    initPairing(mcl::BLS12_381);
    std::string model = argv[1];
    int k = 2;
    std::string in_file = argv[k++], conf_file = argv[k++], net_file;
    if (model == "vgg") net_file = argv[k++];
    int pic_cnt = atoi(argv[k++]);
    uint64_t seed = strtoull(argv[k++], nullptr, 0);
    bool real_gens = false;
    std::string tr_out;
    for (; k < argc; ++k) {
        std::string a = argv[k];
        if (a == "--transcript") tr_out = argv[++k];
        else if (a == "--gens") real_gens = std::string(argv[++k]) == "real";
    }
    if (real_gens) install_real_base_point();
    SeededStream rng(seed);
    rng.install();

    auto t0 = std::chrono::steady_clock::now();
    prover p;
    zkcnn_b200::Transcript tr;
    p.setTranscript(&tr);
    std::unique_ptr<neuralNetwork> nn;
    if (model == "lenet") nn.reset(new lenet(32, 32, 1, pic_cnt, MAX, in_file, conf_file, ""));
    else if (model == "vgg") nn.reset(new vgg(32, 32, 3, pic_cnt, in_file, conf_file, "", net_file));
    else { fprintf(stderr, "unknown model\n"); return 2; }
    nn->create(p, false);
    auto t1 = std::chrono::steady_clock::now();
    bool ok = false;
    try {
        verifier v(&p, p.C);
        ok = v.verify();
    } catch (const std::exception &e) {
        fprintf(stderr, "EXCEPTION: %s\n", e.what());
    }
    auto t2 = std::chrono::steady_clock::now();
    printf("RESULT ok %d n_fr %lu n_g1 %lu bytes %zu fnv %016lx challenges %lu create_s %.3f verify_wall_s %.3f upload_s %.3f launches %lu version \"%s\"\n",
           (int) ok, tr.n_fr, tr.n_g1, tr.bytes.size(), tr.fnv1a(), rng.calls, std::chrono::duration<double>(t1 - t0).count(),
           std::chrono::duration<double>(t2 - t1).count(), p.uploadTime(), p.gpuLaunches(), zk_version());
    printf("TABLE ");
    for (auto &s : output_tb) printf("%s, ", s.c_str());
    puts("");
    if (!tr_out.empty() && !tr.save(tr_out)) { fprintf(stderr, "cannot write %s\n", tr_out.c_str()); return 3; }
    return ok ? 0 : 1;
}
