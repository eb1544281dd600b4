// ref_restir.cpp — TEST INFRASTRUCTURE. C-callable wrapper over the reference's own ReSTIR data structures, compiled for the host in place
// (Shaders/CppCommon/ReSTIRData.h: Reservoir::Update :122-146, Reservoir::UpdateWeight :151-163, CDF::Insert :194-203, CDF::Get /
// BinarySearch :232-302) by oracle/Makefile into oracle/_ref/libref_restir.so. Nothing of the reference is copied. Used by
// tests/golden/make_golden_restir.py to record known answers that pin the oracle's restatement of these functions.
// Asserts are compiled out (NDEBUG): on the device they only exist in debug builds.
#include "shim.h"
#include <cassert>
#include <cstdio>
#include <cstdlib>
using std::isnan; using std::isinf;
#include <Optix/optix.h>              // host API only: the device half of optix.h (inline PTX) must not be pulled in
#include <Cuda/cuda/helpers.h>
#define __CUDACC__ 1
#include "ReSTIRData.h"

extern "C" {

// n sequential Reservoir::Update(sample_k, weights[k], seeds[k]) calls followed by UpdateWeight(); sample_k carries solidAnglePdf = pdfs[k]
// and its own index in radiance.x. out = {weightSum, (float)sampleCount, weight, index of the kept sample, its pdf}; selected[k] = return value.
void ref_kat_reservoir(const float* weights, const unsigned* seeds, const float* pdfs, unsigned n, float* out5, unsigned char* selected)
{
    Reservoir r;
    r.sample.radiance.x = -1.f;
    for (unsigned k = 0; k < n; ++k) {
        LightSample s; s.solidAnglePdf = pdfs[k]; s.radiance.x = (float)k;
        selected[k] = r.Update(s, weights[k], seeds[k]) ? 1 : 0;
    }
    r.UpdateWeight();
    out5[0] = r.weightSum; out5[1] = (float)r.sampleCount; out5[2] = r.weight; out5[3] = r.sample.radiance.x; out5[4] = r.sample.solidAnglePdf;
}

// CDF built by n serial Insert(weight) calls (the accumulated sums are returned in cdf_out), then m Get(value) look-ups
void ref_kat_cdf(const float* weights, unsigned n, const float* values, unsigned m, float* cdf_out, unsigned* index, float* pdf)
{
    CDF* c = static_cast<CDF*>(malloc(sizeof(CDF) + sizeof(float) * (n + 1)));
    c->Reset();
    for (unsigned k = 0; k < n; ++k) c->Insert(weights[k]);
    for (unsigned k = 0; k < n; ++k) cdf_out[k] = c->data[k];
    for (unsigned k = 0; k < m; ++k) c->Get(values[k], index[k], pdf[k]);
    free(c);
}

}
