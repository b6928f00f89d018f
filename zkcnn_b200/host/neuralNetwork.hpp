// Circuit compiler + witness generator of the stand-alone build: turns a CNN description into the layered arithmetic
// circuit (prover::C) and its per-layer values (prover::val) that the GKR prover consumes.
//
// This is the caller-side data format of the hot path, written from scratch but producing,
// gate for gate and value for value, what the reference's neuralNetwork::create does (src/neuralNetwork.cpp:60-142 and
// the layer builders :144-649) -- tests/_cases.py: prove_and_compare checks per-layer hashes of gates, ori_id and values
// against dumps of the compiled reference (tests/golden/*.circuit.txt), up to full-size vgg11 and vgg16.  Public interface = the reference's (src/neuralNetwork.hpp:57-66, src/models.hpp).
#pragma once
#include "prover.hpp"
#include <functional>

enum convType { FFT, NAIVE, NAIVE_FAST };
enum poolType { AVG, MAX, NONE };
enum actType { RELU_ACT };

struct convKernel {
    convType ty;
    i64 channel_out, channel_in, size, stride_bl, padding, weight_start_id, bias_start_id;
    convKernel(convType _ty, i64 _channel_out, i64 _channel_in, i64 _size, i64 _log_stride, i64 _padding)
        : ty(_ty), channel_out(_channel_out), channel_in(_channel_in), size(_size), stride_bl(_log_stride), padding(_padding),
          weight_start_id(0), bias_start_id(0) {}
    convKernel(convType _ty, i64 _channel_out, i64 _channel_in, i64 _size)
        : convKernel(_ty, _channel_out, _channel_in, _size, 0, _size >> 1) {}
};
struct fconKernel {
    i64 channel_out, channel_in, weight_start_id, bias_start_id;
    fconKernel(i64 _channel_out, i64 _channel_in) : channel_out(_channel_out), channel_in(_channel_in), weight_start_id(0), bias_start_id(0) {}
};
struct poolKernel {
    poolType ty;
    i64 size, stride_bl, dcmp_start_id, max_start_id, max_dcmp_start_id;
    poolKernel(poolType _ty, i64 _size, i64 _log_stride) : ty(_ty), size(_size), stride_bl(_log_stride), dcmp_start_id(0), max_start_id(0), max_dcmp_start_id(0) {}
};

// where the image and the weights come from: a whitespace-separated decimal file (the reference's format,
// src/neuralNetwork.cpp:805-897) or an in-memory array (synthetic benchmarks)
class NumberSource {
public:
    virtual ~NumberSource() {}
    virtual double next() = 0;
};

class neuralNetwork {
public:
    explicit neuralNetwork(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, const string &i_filename, const string &c_filename,
                           const string &o_filename);
    virtual ~neuralNetwork() {}

    void create(prover &pr, bool only_compute);

    // ---- additions of the B200 build ----------------------------------------------------------------------------------
    void setInput(std::unique_ptr<NumberSource> src) { in = std::move(src); }
    // number of decimals create() will consume (image + all weights and biases)
    i64 inputCount();
    int hostThreads = 0;          // witness evaluation threads (0 = hardware concurrency)
    vector<int> inferred;         // argmax per picture (what the reference writes to o_file)

    // ---- a new picture for the circuit create() built (device witness generation) --------------------------------
    // create() also records, for the prover (prover::aux_ops), the auxiliary inputs every layer's construction derives from earlier gate
    // values, and the data-dependent quantisation decisions it took: the scale of the picture (calcInputLayer) and of every activation
    // tensor (getNextBit).  Those decide the bit widths of the ReLU / pooling decompositions, i.e. the circuit's STRUCTURE: a new picture
    // can reuse the circuit iff it leads to the same decisions.
    bool deviceWitnessSupported() const { return aux_supported_; }
    // the picture as create() would put it into val[0] (quantised, replicated pic_parallel times); false: its scale differs from the built circuit's
    bool quantizeImage(const double *pixels, size_t n, vector<F> &out) const;
    // ranges[2 * layer], ranges[2 * layer + 1] = largest non-negative value / largest magnitude of a negative value of that layer
    // (zk_witness_generate); true iff every getNextBit decision of create() comes out the same
    bool scalesMatch(const uint64_t *ranges, size_t n_layers) const;
    std::vector<uint8_t> scaleDecisionLayers(size_t n_layers) const;   // flag per layer: a getNextBit decision was taken on its values
    i64 imagePixels() const { return pic_size_x * pic_size_y * pic_channel; }
    // argmax of the output layer (what printInfer does) from its values
    void inferFromOutput(const vector<F> &output);

protected:
    void initParam();
    int getNextBit(int layer_id);
    void refreshConvParam(i64 new_nx, i64 new_ny, const convKernel &conv);
    void calcSizeAfterPool(const poolKernel &p);
    void refreshFCParam(const fconKernel &fc);
    i64 getFFTLen() const;
    i8 getFFTBitLen() const;
    i64 getPoolDecmpSize() const;

    void prepareDecmpBit(i64 layer_id, i64 idx, i64 dcmp_id, i64 bit_shift);
    void prepareFieldBit(const F &data, i64 dcmp_id, i64 bit_shift);
    void prepareSignBit(i64 layer_id, i64 idx, i64 dcmp_id);
    void prepareMax(i64 layer_id, i64 idx, i64 max_id);

    void calcInputLayer(layer &circuit);
    void calcNormalLayer(const layer &circuit, i64 layer_id);
    void calcDotProdLayer(const layer &circuit, i64 layer_id);
    void calcFFTLayer(const layer &circuit, i64 layer_id);

    void inputLayer(layer &circuit);
    void paddingLayer(layer &circuit, i64 &layer_id, i64 first_conv_id);
    void fftLayer(layer &circuit, i64 &layer_id);
    void dotProdLayer(layer &circuit, i64 &layer_id);
    void ifftLayer(layer &circuit, i64 &layer_id);
    void addBiasLayer(layer &circuit, i64 &layer_id, i64 first_bias_id);
    enum class ConvEmit { FUSED, PRODUCTS, SUMS };
    template <class PerOutput, class PerTap> void walkConvolution(PerOutput per_output, PerTap per_tap);
    void naiveConvLayer(layer &circuit, i64 &layer_id, ConvEmit mode, i64 first_conv_id, i64 first_bias_id);
    void reluLayer(layer &circuit, i64 &layer_id, i64 n_activations);
    template <class PerElem, class PerCell> void walkPooling(PerElem per_elem, PerCell per_cell);
    void avgPoolingLayer(layer &circuit, i64 &layer_id);
    void maxPoolingLayer(layeredCircuit &C, i64 &layer_id, i64 first_dcmp_id, i64 first_max_id, i64 first_max_dcmp_id);
    void fullyConnLayer(layer &circuit, i64 &layer_id, i64 first_fc_id, i64 first_bias_id);

    void readBias(i64 first_bias_id);
    void readConvWeight(i64 first_conv_id);
    void readFconWeight(i64 first_fc_id);
    void printInfer(prover &pr);

    vector<vector<convKernel>> conv_section;
    vector<poolKernel> pool;
    poolType pool_ty;
    i64 pool_bl, pool_sz;
    i64 pool_stride_bl, pool_stride;
    i64 pool_layer_cnt, act_layer_cnt, conv_layer_cnt;
    actType act_ty;
    vector<fconKernel> full_conn;

    i64 pic_size_x, pic_size_y, pic_channel, pic_parallel;
    i64 SIZE;
    const i64 NCONV_FAST_SIZE, NCONV_SIZE, FFT_SIZE, AVE_POOL_SIZE, FC_SIZE, RELU_SIZE;
    i64 T;
    const i64 Q = 9;
    i64 Q_MAX;
    const i64 Q_BIT_SIZE = 220;

    i64 nx_in, nx_out, ny_in, ny_out, m, channel_in, channel_out, log_stride, padding;
    i64 new_nx_in, new_ny_in;
    i64 nx_padded_in, ny_padded_in;
    i64 total_in_size, total_para_size, total_relu_in_size, total_ave_in_size, total_max_in_size;
    int x_bit, w_bit, x_next_bit;

    vector<vector<F>>::iterator val;
    vector<F>::iterator two_mul;

    std::unique_ptr<NumberSource> in;
    string o_file;

    // recorded by create()
    struct ScaleDecision { int layer, x_bit, w_bit, next_bit; };
    vector<ScaleDecision> scale_decisions_;
    int image_bit_ = 0;
    bool aux_supported_ = true;
    vector<vector<uint32_t>> *aux_ops_ = nullptr;   // prover::aux_ops
    i64 aux_step_ = 0;
    void recordAux(i64 src_layer, i64 src_idx, i64 dst_idx, u32 kind, i64 bit);
};

// ---- model zoo (src/models.hpp) -------------------------------------------------------------------------------------------
class vgg : public neuralNetwork {
public:
    // network description: integers = conv output channels, 'M' / 'A' = max / average pooling (src/models.cpp:18-35)
    explicit vgg(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, const std::string &i_filename, const string &c_filename,
                 const std::string &o_filename, const std::string &n_filename);
    static std::unique_ptr<vgg> fromDescription(i64 psize, i64 pchannel, i64 pparallel, const std::string &description);
private:
    void configure(std::istream &config_in);
};
class vgg16 : public neuralNetwork {
public:
    explicit vgg16(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, poolType pool_ty, const std::string &i_filename,
                   const string &c_filename, const std::string &o_filename);
};
class vgg11 : public neuralNetwork {
public:
    explicit vgg11(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, poolType pool_ty, const std::string &i_filename,
                   const string &c_filename, const std::string &o_filename);
};
class lenet : public neuralNetwork {
public:
    explicit lenet(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, poolType pool_ty, const std::string &i_filename,
                   const string &c_filename, const std::string &o_filename);
};
class lenetCifar : public neuralNetwork {
public:
    explicit lenetCifar(i64 psize_x, i64 psize_y, i64 pchannel, i64 pparallel, poolType pool_ty, const std::string &i_filename,
                        const string &c_filename, const std::string &o_filename);
};
class ccnn : public neuralNetwork {
public:
    explicit ccnn(i64 psize_x, i64 psize_y, i64 pparallel, i64 pchannel, poolType pool_ty);
};

// host-side helpers shared with the verifier (src/utils.hpp)
long matIdx(long x, long y, long n);
long cubIdx(long x, long y, long z, long n, long m);
long tesIdx(long w, long x, long y, long z, long n, long m, long l);
F getRootOfUnit(int n);
void fft(vector<F> &arr, int logn, bool flag);
