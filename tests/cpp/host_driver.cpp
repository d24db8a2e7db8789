// host_driver.cpp -- exercises the C++ host mirror (gomel_b200/host/gomel.hpp) exactly the way the
// cgo binding would drive the C ABI (one call per API method), and checks every result against the
// float64 CPU oracle.  Test infrastructure: links the oracle; run on the GPU box by
// tests/test_gpu_cpp_host.py.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../gomel_b200/host/gomel.hpp"
#include "../../oracle/gomel_oracle.h"

static double rel_l2(const double* a, const double* b, size_t n)
{
    double num = 0, den = 0;
    for (size_t i = 0; i < n; i++) { num += (a[i] - b[i]) * (a[i] - b[i]); den += b[i] * b[i]; }
    return den > 0 ? std::sqrt(num / den) : std::sqrt(num);
}

int main()
{
    using namespace gomel;
    int bad = 0;
    // deterministic test clip: chirp + tone + LCG noise, 1.1 s
    const long n = 48510;
    std::vector<double> wav(n);
    unsigned long long s = 12345;
    for (long i = 0; i < n; i++) {
        const double t = double(i) / 44100.0;
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        const double noise = (double)(s >> 11) / 9007199254740992.0 - 0.5;
        wav[i] = 0.3 * std::sin(2 * M_PI * (300.0 * t + 2000.0 * t * t)) + 0.2 * std::sin(2 * M_PI * 3100.0 * t) + 0.05 * noise;
    }
    // ---- mel: the cmd/tomel configuration
    mel::Mel* m = mel::NewMel();
    if (m->NumMels != 160 || m->Window != 256 || m->Resolut != 2048) { printf("NewMel defaults wrong\n"); bad++; }
    {   // NewMel defaults (Window 256 / Resolut 2048) run as they are
        orc_config od{};
        od.num_mels = 160; od.num_freqs = 768; od.window = 256; od.resolut = 2048; od.mel_fmin = 0; od.mel_fmax = 8000;
        od.tune_mul = 1; od.tune_add = 0; od.volume_boost = 0; od.gl_iters = 2;
        auto r = m->ToMel(wav);
        if (!r.second.empty()) { printf("ToMel (defaults) error: %s\n", r.second.c_str()); return 1; }
        const long fr = (long)r.first.size() / 160;
        std::vector<double> om(r.first.size() * 2);
        if (orc_to_mel(&od, wav.data(), n, om.data(), (long)om.size()) != fr) { printf("oracle frames differ (defaults)\n"); bad++; }
        std::vector<double> a(om.size()), b(om.size());
        for (size_t i = 0; i < om.size(); i++) { a[i] = std::exp((&r.first[0][0])[i]); b[i] = std::exp(om[i]); }
        const double e = rel_l2(a.data(), b.data(), a.size());
        printf("ToMel    (NewMel defaults) frames=%ld rel-L2(linear)=%.3e\n", fr, e);
        if (!(e < 1e-5)) bad++;
    }
    {   // a geometry outside this build must fail loudly, not fall back
        m->Window = 512; m->Resolut = 1024;
        auto r = m->ToMel(wav);
        if (r.second.empty()) { printf("unsupported config did not fail\n"); bad++; }
    }
    m->NumMels = 192; m->MelFmin = 0; m->MelFmax = 16000; m->Window = 1280; m->Resolut = 4096; m->GriffinLimIterations = 3;
    orc_config oc{};
    oc.num_mels = 192; oc.num_freqs = 768; oc.window = 1280; oc.resolut = 4096; oc.mel_fmin = 0; oc.mel_fmax = 16000;
    oc.tune_mul = 1; oc.tune_add = 0; oc.volume_boost = 0; oc.gl_iters = 3;
    auto rm = m->ToMel(wav);
    if (!rm.second.empty()) { printf("ToMel error: %s\n", rm.second.c_str()); return 1; }
    const long frames = (long)rm.first.size() / 192;
    std::vector<double> omel(rm.first.size() * 2);
    if (orc_to_mel(&oc, wav.data(), n, omel.data(), (long)omel.size()) != frames) { printf("oracle frames differ\n"); bad++; }
    {
        std::vector<double> a(omel.size()), b(omel.size());
        for (size_t i = 0; i < omel.size(); i++) { a[i] = std::exp((&rm.first[0][0])[i]); b[i] = std::exp(omel[i]); }
        const double e = rel_l2(a.data(), b.data(), a.size());
        printf("ToMel    frames=%ld rel-L2(linear)=%.3e\n", frames, e);
        if (!(e < 1e-5)) bad++;
    }
    const long ola = 4096 + (frames - 1) * 1280;
    std::vector<double> init(ola);
    for (auto& v : init) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; v = (double)(s >> 11) / 9007199254740992.0; }
    m->InitSignal = init;
    std::vector<Pair> spec((size_t)frames * 192);
    for (size_t i = 0; i < spec.size(); i++) { spec[i][0] = omel[2 * i]; spec[i][1] = omel[2 * i + 1]; }
    auto rw = m->FromMel(spec);
    if (!rw.second.empty()) { printf("FromMel error: %s\n", rw.second.c_str()); return 1; }
    {
        std::vector<double> ow(ola), omel2 = omel;
        orc_from_mel(&oc, omel2.data(), frames * 192, init.data(), ow.data(), ola);
        const double e = rel_l2(rw.first.data(), ow.data(), (size_t)ola);
        const double side = std::fabs(spec[5][0] - std::exp(omel[10]));
        printf("FromMel  GL-3 rel-L2=%.3e  in-place exp side effect err=%.1e\n", e, side);
        if (!(e < 1e-4) || !(side < 1e-12)) bad++;
    }
    {
        std::vector<Pair> bad_len(192 * 2 + 48);
        auto r = m->FromMel(bad_len);
        if (r.second.empty()) { printf("ragged input did not fail\n"); bad++; }
    }
    {
        std::vector<Pair> mp((size_t)frames * 192);
        for (size_t i = 0; i < mp.size(); i++) { mp[i][0] = omel[2 * i]; mp[i][1] = omel[2 * i + 1]; }
        auto img = m->Image(mp);
        std::vector<unsigned short> oimg(mp.size());
        orc_mel_dumpbuffer(omel.data(), (long)mp.size(), 192, oimg.data());
        size_t diff = 0;
        for (size_t i = 0; i < oimg.size(); i++) diff += img[i] != oimg[i];
        printf("Image    %zu entries, %zu differ\n", oimg.size(), diff);
        if (diff) bad++;
    }
    // ---- phase: NewPhase defaults are the supported configuration
    phase::Phase* p = phase::NewPhase();
    auto rp = p->ToPhase(wav);
    if (!rp.second.empty()) { printf("ToPhase error: %s\n", rp.second.c_str()); return 1; }
    std::vector<double> ophase(rp.first.size() * 2);
    orc_to_phase(&oc, wav.data(), n, ophase.data(), (long)ophase.size());
    {
        const double e = rel_l2(&rp.first[0][0], ophase.data(), ophase.size());
        printf("ToPhase  rel-L2=%.3e\n", e);
        if (!(e < 1e-5)) bad++;
    }
    p->VolumeBoost = 1.666;
    oc.volume_boost = 1.666;
    std::vector<Pair> ps(rp.first.size());
    for (size_t i = 0; i < ps.size(); i++) { ps[i][0] = ophase[2 * i]; ps[i][1] = ophase[2 * i + 1]; }
    auto rf = p->FromPhase(ps);
    if (!rf.second.empty()) { printf("FromPhase error: %s\n", rf.second.c_str()); return 1; }
    {
        std::vector<double> of(rf.first.size());
        orc_from_phase(&oc, ophase.data(), (long)ps.size(), of.data(), (long)of.size());
        const double e = rel_l2(rf.first.data(), of.data(), of.size());
        printf("FromPhase rel-L2=%.3e\n", e);
        if (!(e < 1e-5)) bad++;
    }
    printf(bad ? "CPP_HOST_FAIL %d\n" : "CPP_HOST_OK\n", bad);
    return bad ? 1 : 0;
}
