// ldpcsim_main.cpp -- host mains that keep the reference binaries' entry points.
//
// One executable; the variant is picked from argv[0]'s basename (bin/decodeMinSum, bin/decodeSMNGDBF,
// ... are links to bin/ldpcsim) or from a leading `ldpcsim <variant>` argument.  For every reference
// binary (C_implementations/Makefile:24-71, plus NGDBFhw) this keeps
//   * the positional command line, whose arity depends on the variant's macro set
//     (src/decodeMinSum.cpp:74-143, src/decodeGDBF.cpp:88-164, src/RNGDBF.cpp:82-156,
//      src/decodeDDBMP.cpp, src/decodeBP.cpp:60-103, src/NGDBFhw.cpp:478-511),
//   * a wrong argument count printing the usage line and returning 0,
//   * the parameter echo, the "Incremental result" / "Final result" phrases,
//   * the stop rule  while (errors < 200 || wordErrors < W)  (src/decodeMinSum.cpp:189,
//     src/decodeBP.cpp:145-151, src/decodeGDBF.cpp:222-226), polled per launch batch,
//   * one appended tab-separated result line per run with the variant's own columns
//     (src/decodeMinSum.cpp:313-329, src/decodeGDBF.cpp:425-453, src/RNGDBF.cpp:459-483,
//      src/decodeDDBMP.cpp:255-265, src/NGDBFhw.cpp:449-469 incl. <log>_<SNR>_itdist.dat).
// The Monte-Carlo loop itself runs in ldpc_gpu_simulate() (include/ldpc_gpu.h); nothing is decoded
// on the CPU.  Knobs the reference does not have come from the environment so that the positional
// interface stays untouched:
//   LDPC_SEED       Philox seed (default time(0), like ran_seed(time(0)), src/decodeMinSum.cpp:187)
//   LDPC_FRAMES     fixed frame count instead of the stop rule
//   LDPC_PRECISION  f64 (default: the reference's arithmetic) | f32 | f16x2 (min-sum family: binary16 messages; results identical to
//                   f64 when the macro set is an exact lattice -- stderr says which kernel runs, include/ldpc_gpu.h LDPC_GPU_PREC_F16X2)
//   LDPC_DEVICES    comma list of CUDA ordinals; frames are sharded by frame-id range, one host
//                   thread per GPU (default "0")
//   LDPC_POLL       frames per launch between stop-rule polls
//   LDPC_RANDOM_CODEWORDS=n   when no codeword file is given: send n random codewords of H (host GF(2) encoder)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "ldpc_gpu.h"

using namespace std;

namespace {

struct Variant { const char *name; int kind; unsigned flags; };
const unsigned QS = LDPC_GPU_F_QUANTIZE_SAMPLES, SS = LDPC_GPU_F_SATURATE_SAMPLES, NM = LDPC_GPU_F_NORMALIZED_MS,
               OM = LDPC_GPU_F_OFFSET_MS, SEQ = LDPC_GPU_F_SEQUENTIALMODE, MSW = LDPC_GPU_F_MODESWITCHING,
               AN = LDPC_GPU_F_ADDNOISE, WS = LDPC_GPU_F_WEIGHTSYNDROMES, OS = LDPC_GPU_F_OUTPUTSMOOTHING,
               TA = LDPC_GPU_F_THRESHOLDADAPTATION, QP = LDPC_GPU_F_QUANTIZEPROBABILITIES, RD = LDPC_GPU_F_REDECODE,
               UN = LDPC_GPU_F_UNIFORMNOISE;
const Variant kVariants[] = {
    { "decodeMinSum", LDPC_GPU_KIND_MINSUM, 0 },                                  // Makefile:58
    { "decodeOffsetMinSum", LDPC_GPU_KIND_MINSUM, QS | OM },                      // :61
    { "decodeNormalizedMinSum", LDPC_GPU_KIND_MINSUM, QS | NM },                  // :64
    { "decodeBP", LDPC_GPU_KIND_BP, 0 },                                          // :67
    { "decodeDDBMP", LDPC_GPU_KIND_DDBMP, 0 },                                    // :70
    { "decodeGDBF", LDPC_GPU_KIND_GDBF, 0 },                                      // plain parallel GDBF (no goal)
    { "decodeMGDBF", LDPC_GPU_KIND_GDBF, MSW },                                   // :24
    { "decodeSGDBF", LDPC_GPU_KIND_GDBF, SEQ },                                   // :27
    { "decodeStochasticNGDBF", LDPC_GPU_KIND_GDBF, QS | QP | WS | SS },           // :30
    { "decodeMNGDBF", LDPC_GPU_KIND_GDBF, AN | TA | WS | SS },                    // :33
    { "decodeSMNGDBF", LDPC_GPU_KIND_GDBF, AN | TA | WS | OS | SS },              // :36
    { "decodeUniformSMNGDBF", LDPC_GPU_KIND_GDBF, AN | UN | TA | WS | OS | SS },  // scripts/ngdbf_example_uniform_*.sh
    { "decodeRSMNGDBF", LDPC_GPU_KIND_GDBF, RD | AN | TA | WS | OS | SS },        // :45
    { "decodeSMGDBF", LDPC_GPU_KIND_GDBF, OS },                                   // :49
    { "decodeSATGDBF", LDPC_GPU_KIND_GDBF, TA | OS },                             // :52
    { "decodeATGDBF", LDPC_GPU_KIND_GDBF, TA },                                   // :55
    { "NGDBFhw", LDPC_GPU_KIND_NGDBF_HW, 0 },                                     // scripts/demo_NGDBFhw_802_3.sh:24
};

const Variant *find_variant(const string &n)
{
    for (const Variant &v : kVariants) if (n == v.name) return &v;
    return nullptr;
}

// positional argument names, in the reference's order
vector<string> usage_of(const Variant &v)
{
    vector<string> a;
    const unsigned f = v.flags;
    if (v.kind == LDPC_GPU_KIND_NGDBF_HW) { a = { "alist", "SNR", "numFrames", "seed", "logfilename" }; }
    else {
        a = { "alist", "R", "SNR", "T" };
        if (v.kind == LDPC_GPU_KIND_MINSUM) {
            if (f & (QS | SS)) a.push_back("Ymax");
            if (f & QS) a.push_back("Q");
            if (f & NM) a.push_back("alpha");
            if (f & OM) a.push_back("delta");
            a.push_back("logfilename");
        } else if (v.kind == LDPC_GPU_KIND_BP) a.push_back("logfilename");
        else if (v.kind == LDPC_GPU_KIND_DDBMP) { a.push_back("Ymax"); a.push_back("Q"); a.push_back("logfilename"); }
        else {
            a.push_back("theta"); a.push_back("logfilename");
            if (f & (AN | QP)) a.push_back("noiseScale");
            if ((f & QS) && !(f & RD)) a.push_back("NQ");
            if (f & TA) a.push_back("lambda");
            if (f & WS) a.push_back("alpha");
            if (f & OS) a.push_back("windowsize");
            if (f & SS) a.push_back("Ymax");
            if (f & RD) a.push_back("maxphase");
        }
    }
    a.push_back("[codeword filename]");
    return a;
}

bool load_codewords(const string &path, int N, vector<uint8_t> &out, long &n)
{   // data.enc: one codeword per line, '0'/'1' (src/decodeMinSum.cpp:195-211)
    ifstream f(path.c_str());
    if (!f) return false;
    string s; n = 0;
    while (getline(f, s)) {
        if ((int)s.size() < N) continue;
        for (int i = 0; i < N; i++) {
            if (s[i] == '1') out.push_back(1);
            else { if (s[i] != '0') cout << "Got an invalid symbol at index " << i << endl; out.push_back(0); }
        }
        n++;
    }
    return n > 0;
}

void print_hist(const vector<int64_t> &h)
{   // printHistogram, src/decodeMinSum.cpp:373-380
    for (size_t i = 0; i < h.size(); i++) if (h[i] > 0) cout << i + 1 << ":\t" << h[i] << endl;
}

struct Shard {
    ldpc_gpu_decoder *dec = nullptr; int device = 0;
    ldpc_gpu_counters cnt; vector<int64_t> ew, ith, ph; int rc = 0; string err;
};

// redecodeStatistics: src/redecodeStatistics.cpp with the macro set of the noisy smoothed decoder
// (addNoise thresholdAdaptation weightSyndromes outputSmoothing saturateSamples).  Every frame is decoded NR times from
// the same channel samples with fresh perturbation noise; one row of NR error weights per frame is appended to the log
// (src/redecodeStatistics.cpp:392-395, fprintVector :615-621).
int redecode_main(int argc, char *argv[])
{
    const vector<string> args = { "alist", "R", "SNR", "T", "NR", "NF", "theta", "logfilename", "noiseScale", "lambda", "alpha",
                                  "windowsize", "Ymax", "[codeword filename]" };
    if ((argc != (int)args.size()) && (argc != (int)args.size() + 1)) {
        cout << "Usage: " << argv[0];
        for (size_t i = 0; i < args.size(); i++) cout << " " << args[i];
        cout << "\n";
        return 0;
    }
    ldpc_gpu_decoder_cfg cfg;
    ldpc_gpu_decoder_cfg_default(LDPC_GPU_KIND_GDBF, &cfg);
    cfg.flags = RD | AN | TA | WS | OS | SS;                      // RD with maxphase = 1: the file's weight alpha*Ymax/dv (:543-546), one phase per decode
    cfg.maxphase = 1;
    const char *pe = getenv("LDPC_PRECISION");
    cfg.precision = (pe && string(pe) == "f32") ? LDPC_GPU_PREC_F32 : LDPC_GPU_PREC_F64;
    int idx = 1;
    ldpc_gpu_code *code = nullptr;
    if (ldpc_gpu_code_load_alist(argv[idx++], &code)) { cerr << "alist: " << ldpc_gpu_last_error() << endl; return 1; }
    int N, M, E, dv, dc;
    ldpc_gpu_code_dims(code, &N, &M, &E, &dv, &dc);
    cout << "PARAMETERS: \n alist = \t" << argv[1] << endl;
    const double R = atof(argv[idx++]);   cout << " R = \t" << R << endl;
    const double SNR = atof(argv[idx++]); cout << " SNR = \t" << SNR << endl;
    cfg.num_iterations = atoi(argv[idx++]); cout << " T = \t" << cfg.num_iterations << endl;
    const int NR = atoi(argv[idx++]);     cout << " NR = \t" << NR << endl;
    const long NF = atol(argv[idx++]);    cout << " NF = \t" << NF << endl;
    cfg.theta = atof(argv[idx++]);        cout << " theta = \t" << cfg.theta << endl;
    const string logfilename = argv[idx++]; cout << " log = \t" << logfilename << endl;
    cfg.noiseScale = atof(argv[idx++]);   cout << " noiseScale = \t" << cfg.noiseScale << endl;
    cfg.lambda = atof(argv[idx++]);       cout << " lambda = \t" << cfg.lambda << endl;
    cfg.alpha = atof(argv[idx++]);        cout << " alpha = \t" << cfg.alpha << endl;
    cfg.windowsize = atoi(argv[idx++]);   cout << "windowsize = \t" << cfg.windowsize << endl;
    cfg.Ymax = atof(argv[idx++]);         cout << " Ymax = \t" << cfg.Ymax << endl;
    vector<uint8_t> cw; long n_cw = 0;
    if (argc == (int)args.size() + 1) {
        cout << "\nUsing codewords from " << argv[idx] << endl;
        if (!load_codewords(argv[idx], N, cw, n_cw)) { cerr << "cannot read codewords from " << argv[idx] << endl; return 1; }
    } else cout << "\nUsing all-zero sequence.\n";
    if (NR < 1 || NF < 0) { cerr << "NR must be >= 1 and NF >= 0" << endl; return 1; }
    vector<int> devs;
    { const char *e = getenv("LDPC_DEVICES"); string sdev = e ? e : "0"; stringstream ss(sdev); string t;
      while (getline(ss, t, ',')) if (!t.empty()) devs.push_back(atoi(t.c_str())); }
    if (ldpc_gpu_init(devs.data(), (int)devs.size())) { cerr << "ldpc_gpu_init: " << ldpc_gpu_last_error() << endl; return 1; }
    const char *es = getenv("LDPC_SEED");
    const uint64_t philox_seed = es ? strtoull(es, 0, 10) : (uint64_t)time(0);
    const ldpc_gpu_channel ch = { SNR, R };
    const size_t G = devs.size();
    vector<vector<int32_t>> out(G);
    vector<ldpc_gpu_counters> cnt(G);
    vector<int> rcs(G, 0); vector<string> errs(G);
    vector<thread> th;
    for (size_t g = 0; g < G; g++) {                              // contiguous frame-id ranges, one per GPU
        const long long begin = (long long)NF * (long long)g / (long long)G, end = (long long)NF * (long long)(g + 1) / (long long)G;
        out[g].assign((size_t)(end - begin) * NR, 0);
        memset(&cnt[g], 0, sizeof cnt[g]);
        th.emplace_back([&, g, begin, end]() {
            ldpc_gpu_decoder *dec = nullptr;
            if (ldpc_gpu_decoder_create(code, &cfg, devs[g], &dec)) { rcs[g] = 1; errs[g] = ldpc_gpu_last_error(); return; }
            if (n_cw) ldpc_gpu_decoder_set_codewords(dec, cw.data(), n_cw);
            ldpc_gpu_sim_args a = { philox_seed, begin, end - begin, 0, 0, 0 };
            if (ldpc_gpu_redecode_stats(dec, &ch, &a, NR, out[g].data(), &cnt[g])) { rcs[g] = 1; errs[g] = ldpc_gpu_last_error(); }
            ldpc_gpu_decoder_destroy(dec);
        });
    }
    for (thread &t : th) t.join();
    for (size_t g = 0; g < G; g++) if (rcs[g]) { cerr << "ldpc_gpu_redecode_stats: " << errs[g] << endl; return 1; }
    ofstream of(logfilename.c_str(), ios::app);
    long long errors = 0, totalBits = 0, totalIterations = 0, uncodedErrors = 0;
    for (size_t g = 0; g < G; g++) {
        const size_t rows = out[g].size() / (size_t)NR;
        for (size_t fi = 0; fi < rows; fi++) { for (int r = 0; r < NR; r++) of << out[g][fi * NR + r] << "\t"; of << endl; }
        errors += cnt[g].errors; totalIterations += cnt[g].totalIterations;
        uncodedErrors += cnt[g].uncodedErrors / NR; totalBits += cnt[g].totalBits / NR;    // counted once per frame by the reference
    }
    of.close();
    cout << "\nFinal result: " << errors << " bit errs in " << NF << " words, BER=" << (totalBits ? (double)errors / totalBits : 0.0)
         << ". Average iterations = " << (NF ? (double)totalIterations / ((double)NF * NR) : 0.0) << ". Uncoded errors = " << uncodedErrors
         << ", uncBER=" << (totalBits ? (double)uncodedErrors / totalBits : 0.0) << endl;
    ldpc_gpu_code_destroy(code);
    return 0;
}

} // namespace

int main(int argc, char *argv[])
{
    string prog = argv[0];
    size_t slash = prog.find_last_of('/');
    string base = slash == string::npos ? prog : prog.substr(slash + 1);
    if (base == "redecodeStatistics") return redecode_main(argc, argv);
    if (argc >= 2 && string(argv[1]) == "redecodeStatistics") return redecode_main(argc - 1, argv + 1);
    const Variant *v = find_variant(base);
    if (!v && argc >= 2 && (v = find_variant(argv[1]))) { argv++; argc--; }
    if (!v) {
        cout << "Usage: ldpcsim <variant> <the variant's positional arguments>\nvariants:";
        for (const Variant &x : kVariants) cout << " " << x.name;
        cout << "\n";
        return 0;
    }
    const vector<string> args = usage_of(*v);
    if ((argc != (int)args.size()) && (argc != (int)args.size() + 1)) {
        cout << "Usage: " << argv[0];
        for (size_t i = 0; i < args.size(); i++) cout << " " << args[i];
        cout << "\n";
        return 0;
    }
    const unsigned f = v->flags;
    const bool hw = v->kind == LDPC_GPU_KIND_NGDBF_HW, gdbf = v->kind == LDPC_GPU_KIND_GDBF;
    ldpc_gpu_decoder_cfg cfg;
    ldpc_gpu_decoder_cfg_default(v->kind, &cfg);
    cfg.flags = f;
    const char *pe = getenv("LDPC_PRECISION");
    cfg.precision = (pe && string(pe) == "f32") ? LDPC_GPU_PREC_F32 : (pe && string(pe) == "f16x2") ? LDPC_GPU_PREC_F16X2 : LDPC_GPU_PREC_F64;

    // ---- parse, echoing like the reference --------------------------------------------------
    int idx = 1;
    const char *alist = argv[idx++];
    ldpc_gpu_code *code = nullptr;
    if (ldpc_gpu_code_load_alist(alist, &code)) { cerr << "alist: " << ldpc_gpu_last_error() << endl; return 1; }
    int N, M, E, dv, dc;
    ldpc_gpu_code_dims(code, &N, &M, &E, &dv, &dc);
    cout << "PARAMETERS: \n alist = \t" << argv[1] << endl;
    double R = 0.8413, SNR = 0; long numFrames = 0; long seed = 0; string logfilename; double Ymax_echo = cfg.Ymax;
    if (hw) {
        SNR = atof(argv[idx++]);   cout << " SNR = \t" << SNR << endl;
        numFrames = atoi(argv[idx++]); cout << "Simulating for " << numFrames << " frames." << endl;
        seed = atoi(argv[idx++]);  cout << "Using random seed " << seed << endl;
        logfilename = argv[idx++]; cout << " log = \t" << logfilename << endl;
    } else {
        R = atof(argv[idx++]);   cout << " R = \t" << R << endl;
        SNR = atof(argv[idx++]); cout << " SNR = \t" << SNR << endl;
        cfg.num_iterations = atoi(argv[idx++]); cout << " T = \t" << cfg.num_iterations << endl;
        if (v->kind == LDPC_GPU_KIND_MINSUM) {
            if (f & SS) { cfg.Ymax = atof(argv[idx++]); cout << "Applying sample clipping with Ymax = +/-" << cfg.Ymax << endl; }
            if (f & QS) { cfg.Ymax = atof(argv[idx++]); cfg.Q = atoi(argv[idx++]);
                          cout << "Applying sample quantization with Ymax = +/-" << cfg.Ymax << " on " << cfg.Q << " bits with "
                               << pow(2.0, cfg.Q) - 1 << " non-zero levels." << endl; }
            if (f & NM) { cfg.alpha = atof(argv[idx++]); cout << "Using normalization with alpha=" << cfg.alpha << endl; }
            if (f & OM) { cfg.delta = atof(argv[idx++]); cout << "Using offset MS with delta=" << cfg.delta << endl; }
            logfilename = argv[idx++]; cout << " log = \t" << logfilename << endl;
        } else if (v->kind == LDPC_GPU_KIND_BP) { logfilename = argv[idx++]; cout << " log = \t" << logfilename << endl; }
        else if (v->kind == LDPC_GPU_KIND_DDBMP) {
            cfg.Ymax = atof(argv[idx++]); cfg.Q = atoi(argv[idx++]);
            cout << " Ymax = \t" << cfg.Ymax << endl << "Q = \t" << cfg.Q << endl;   // src/decodeDDBMP.cpp:88-90
            logfilename = argv[idx++]; cout << " log = \t" << logfilename << endl;
        } else {
            cfg.theta = atof(argv[idx++]); cout << " theta = \t" << cfg.theta << endl;
            logfilename = argv[idx++];     cout << " log = \t" << logfilename << endl;
            if (f & (AN | QP)) { cfg.noiseScale = atof(argv[idx++]); cout << " noiseScale = \t" << cfg.noiseScale << endl; }
            if ((f & QS) && !(f & RD)) { cfg.NQ = atoi(argv[idx++]); cout << " NQ = \t" << cfg.NQ << endl; }
            if (f & TA) { cfg.lambda = atof(argv[idx++]); cout << " lambda = \t" << cfg.lambda << endl; }
            if (f & WS) { cfg.alpha = atof(argv[idx++]); cout << " alpha = \t" << cfg.alpha << endl; }
            if (f & OS) { cfg.windowsize = atoi(argv[idx++]); cout << "windowsize = \t" << cfg.windowsize << endl; }
            if (f & SS) { cfg.Ymax = atof(argv[idx++]); cout << " Ymax = \t" << cfg.Ymax << endl; }
            if (f & RD) { cfg.maxphase = atoi(argv[idx++]); cout << " maxphase = \t" << cfg.maxphase << endl; }
        }
        Ymax_echo = cfg.Ymax;
    }
    vector<uint8_t> cw; long n_cw = 0;
    if (argc == (int)args.size() + 1) {
        cout << "\nUsing codewords from " << argv[idx] << endl;
        if (!load_codewords(argv[idx], N, cw, n_cw)) { cerr << "cannot read codewords from " << argv[idx] << endl; return 1; }
    } else if (const char *rc = getenv("LDPC_RANDOM_CODEWORDS")) {
        // no codeword file for this code (e.g. codes/802_3/data_c.enc is missing upstream): encode random ones
        n_cw = atol(rc) > 0 ? atol(rc) : 500;
        cw.resize((size_t)n_cw * N);
        int rank = 0;
        if (ldpc_gpu_code_random_codewords(code, 20261018ull, n_cw, cw.data(), &rank)) { cerr << "encoder: " << ldpc_gpu_last_error() << endl; return 1; }
        cout << "\nUsing " << n_cw << " random codewords (rank(H) = " << rank << ").\n";
    } else cout << "\nUsing all-zero sequence.\n";

    const double N0 = pow(10.0, -SNR / 10.0) / R;
    const double sigma = sqrt(N0 / 2.0);
    const char *what = v->kind == LDPC_GPU_KIND_MINSUM ? "Min-Sum" : v->kind == LDPC_GPU_KIND_BP ? "BP"
                     : v->kind == LDPC_GPU_KIND_DDBMP ? "DD-BMP" : "GDBF";
    cout << "Simulating " << what << " decoding on code with N=" << N << ", M=" << M << ", R=" << R << ", dv=" << dv << ", dc=" << dc << endl;
    cout << "\nParameters are:\n\tSNR\t" << SNR << "\n\tN0\t" << N0 << "\n\tsigma\t" << sigma << endl;

    // ---- devices ---------------------------------------------------------------------------------
    vector<int> devs;
    { const char *e = getenv("LDPC_DEVICES"); string s = e ? e : "0"; stringstream ss(s); string t;
      while (getline(ss, t, ',')) if (!t.empty()) devs.push_back(atoi(t.c_str())); }
    if (ldpc_gpu_init(devs.data(), (int)devs.size())) { cerr << "ldpc_gpu_init: " << ldpc_gpu_last_error() << endl; return 1; }
    const int T = cfg.num_iterations;
    const int ith_len = ldpc_gpu_iter_hist_len(&cfg);
    vector<Shard> sh(devs.size());
    for (size_t g = 0; g < devs.size(); g++) {
        sh[g].device = devs[g];
        if (ldpc_gpu_decoder_create(code, &cfg, devs[g], &sh[g].dec)) { cerr << "decoder: " << ldpc_gpu_last_error() << endl; return 1; }
        if (g == 0 && cfg.precision == LDPC_GPU_PREC_F16X2) {
            int32_t exact = 0; ldpc_gpu_decoder_stats(sh[g].dec, nullptr, &exact);
            cerr << (exact ? "f16x2: exact lattice, results identical to LDPC_PRECISION=f64" : "f16x2: NOT an exact lattice, binary16 messages clamped (labelled throughput arithmetic)") << endl;
        }
        if (n_cw) ldpc_gpu_decoder_set_codewords(sh[g].dec, cw.data(), n_cw);
        sh[g].ew.assign(N, 0); sh[g].ith.assign(ith_len, 0); sh[g].ph.assign(cfg.maxphase > 0 ? cfg.maxphase : 1, 0);
    }

    // ---- main test loop ---------------------------------------------------------------------------
    int minWordErrors = 40;                                       // MS, DD-BMP (src/decodeMinSum.cpp:189)
    if (v->kind == LDPC_GPU_KIND_BP || gdbf) { minWordErrors = 20; if (N > 10000) minWordErrors = 10; if (N > 50000) minWordErrors = 5; }
    const char *es = getenv("LDPC_SEED"), *ef = getenv("LDPC_FRAMES"), *ep = getenv("LDPC_POLL");
    const uint64_t philox_seed = hw ? (uint64_t)seed : (es ? strtoull(es, 0, 10) : (uint64_t)time(0));
    const long long fixed = hw ? numFrames : (ef ? atoll(ef) : 0);
    int grid = 0; ldpc_gpu_decoder_geometry(sh[0].dec, &grid, 0, 0, 0);
    long long poll = ep ? atoll(ep) : (fixed ? (1ll << 22) : (long long)grid * 4);
    if (poll < 1) poll = 1;

    ldpc_gpu_counters tot; memset(&tot, 0, sizeof tot);
    vector<int64_t> ew(N, 0), ith(ith_len, 0), ph(cfg.maxphase > 0 ? cfg.maxphase : 1, 0);
    long long next_frame = 0;
    const ldpc_gpu_channel ch = { SNR, R };
    for (;;) {
        if (fixed) { if (tot.totalWords >= fixed) break; }
        else if (!((tot.errors < 200) || (tot.wordErrors < minWordErrors))) break;
        // one poll: every GPU takes the next frame-id range
        vector<thread> th;
        for (size_t g = 0; g < sh.size(); g++) {
            long long nf = poll;
            if (fixed) nf = min<long long>(nf, fixed - tot.totalWords - (long long)g * poll);
            if (nf <= 0) { sh[g].rc = 0; memset(&sh[g].cnt, 0, sizeof sh[g].cnt); continue; }
            const long long begin = next_frame; next_frame += nf;
            th.emplace_back([&, g, begin, nf]() {
                Shard &s = sh[g];
                memset(&s.cnt, 0, sizeof s.cnt);
                fill(s.ew.begin(), s.ew.end(), 0); fill(s.ith.begin(), s.ith.end(), 0); fill(s.ph.begin(), s.ph.end(), 0);
                s.cnt.error_weight_hist = s.ew.data(); s.cnt.iter_hist = s.ith.data(); s.cnt.phase_hist = s.ph.data();
                ldpc_gpu_sim_args a = { philox_seed, begin, nf, 0, 0, 0 };
                s.rc = ldpc_gpu_simulate(s.dec, &ch, &a, &s.cnt);
                if (s.rc) s.err = ldpc_gpu_last_error();
            });
        }
        for (thread &t : th) t.join();
        for (Shard &s : sh) {
            if (s.rc) { cerr << "ldpc_gpu_simulate: " << s.err << endl; return 1; }
            tot.errors += s.cnt.errors; tot.uncodedErrors += s.cnt.uncodedErrors; tot.totalBits += s.cnt.totalBits;
            tot.totalWords += s.cnt.totalWords; tot.wordErrors += s.cnt.wordErrors; tot.totalIterations += s.cnt.totalIterations;
            tot.smoothingUsed += s.cnt.smoothingUsed; tot.undetectedWords += s.cnt.undetectedWords;
            if (s.cnt.totalWords) {
                for (int i = 0; i < N; i++) ew[i] += s.ew[i];
                for (int i = 0; i < ith_len; i++) ith[i] += s.ith[i];
                for (size_t i = 0; i < ph.size(); i++) ph[i] += s.ph[i];
            }
        }
        cout << "\nIncremental result: " << tot.errors << " bit errs in " << tot.totalWords << " words, BER=" << (double)tot.errors / tot.totalBits
             << ". Average iterations = " << (double)tot.totalIterations / tot.totalWords << ". Word error=" << tot.wordErrors
             << ". Uncoded errors = " << tot.uncodedErrors << ", uncBER=" << (double)tot.uncodedErrors / tot.totalBits
             << "\nError weights:\n";
        print_hist(ew);
        if (f & RD) { cout << "Phase histogram:\n"; print_hist(ph); }
    }

    // ---- final report + appended TSV line ---------------------------------------------------------
    const long errors = tot.errors, totalBits = tot.totalBits, totalWords = tot.totalWords, wordErrors = tot.wordErrors,
               totalIterations = tot.totalIterations, uncodedErrors = tot.uncodedErrors, smoothingUsed = tot.smoothingUsed;
    cout << "\nFinal result: " << errors << " bit errs in " << totalWords << " words, BER=" << (double)errors / totalBits
         << ". Average iterations = " << (double)totalIterations / totalWords << ". Uncoded errors = " << uncodedErrors
         << ", uncBER=" << (double)uncodedErrors / totalBits << endl;
    ofstream of(logfilename.c_str(), ios::app);
    const char tab = '\t';
    if (hw) {                                                     // src/NGDBFhw.cpp:449-459
        of << SNR << tab << errors << tab << wordErrors << tab << (double)errors / totalBits << tab << (double)totalIterations / totalWords << tab
           << (double)wordErrors / totalWords << tab << totalBits << tab << totalWords << tab << cfg.num_iterations << tab << cfg.theta0 << tab;
        of << cfg.noiseScale << tab; of << cfg.w << tab; of << cfg.Ymax << tab << 5 << tab; of << cfg.maxphase << tab << seed; of << endl;
        stringstream ss; ss << logfilename << "_" << SNR << "_itdist.dat";      // :461-469: P(completion time >= idx)
        ofstream ofit(ss.str().c_str(), ios::trunc);
        long long tail = totalWords;
        for (int i = 0; i < cfg.num_iterations; i++) { ofit << i << "\t" << (double)tail / totalWords << "\n"; tail -= ith[i]; }
    } else {
        of << SNR << tab << (double)errors / totalBits << tab << (double)totalIterations / totalWords << tab << (double)wordErrors / totalWords << tab;
        if (gdbf) {                                               // src/decodeGDBF.cpp:425-453 / src/RNGDBF.cpp:459-483
            of << totalBits << tab << totalWords << tab << cfg.num_iterations << tab << cfg.theta << tab;
            if (f & (AN | QP)) of << cfg.noiseScale << tab;
            if ((f & QS) && !(f & RD)) of << cfg.NQ << tab;
            if (f & TA) of << cfg.lambda << tab;
            if (f & WS) of << cfg.alpha << tab;
            if (f & OS) { of << smoothingUsed << tab << (double)smoothingUsed / totalWords << tab; of << cfg.windowsize << tab; }
            if (f & SS) of << cfg.Ymax << tab;
            if (f & RD) of << cfg.maxphase << tab;
        } else {
            of << cfg.num_iterations << tab;
            if (v->kind == LDPC_GPU_KIND_MINSUM) {                // src/decodeMinSum.cpp:313-329
                if (f & (SS | QS)) of << Ymax_echo << tab;
                if (f & NM) of << cfg.alpha << tab;
                if (f & OM) of << cfg.delta << tab;
            } else if (v->kind == LDPC_GPU_KIND_DDBMP) of << cfg.Ymax << tab << cfg.Q << tab;   // src/decodeDDBMP.cpp:255-265
        }
        of << argv[1] << endl;
    }
    of.close();
    for (Shard &s : sh) ldpc_gpu_decoder_destroy(s.dec);
    ldpc_gpu_code_destroy(code);
    return 0;
}
