// Force-included (g++ -include) when the REFERENCE's own translation units (verifier.cpp, neuralNetwork.cpp,
// models.cpp, polyVerifier.cpp, ...) are compiled against the zkcnn_b200 prover: it declares `class prover` and
// `class hyrax_bls12_381::polyProver` with the reference's interface and defines the reference headers' include
// guards, so src/prover.hpp and hyrax/src/polyProver.hpp of the reference are skipped.  See INTEGRATION.md.
#pragma once
#ifndef ZKCNN_DROPIN
#define ZKCNN_DROPIN 1
#endif
#ifndef ZKCNN_DROPIN_CPU_HYRAX
#include "polyProver.hpp"
#endif
#include "prover.hpp"
