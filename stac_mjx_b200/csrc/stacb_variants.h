// Kernel variants: (coords per lane, active bodies per lane, all bodies per lane, sites per lane, joint slots per body).
// The dispatcher picks the first variant that fits the model; build.sh compiles one TU per entry.
#pragma once
#define STACB_VARIANTS(X) X(1, 1, 1, 1, 1) X(3, 1, 3, 1, 3) X(2, 2, 3, 1, 2) X(4, 3, 4, 2, 3) X(8, 6, 7, 2, 1)
// Register-resident hinge-tree solver (stacb_fast.cuh): (joint slots per body, pointer-jumping rounds, all bodies per lane).
// Surplus slots / rounds are exact no-ops; the oracle (mode 2) picks its variant by the same first-fit rule.
#define STACB_FAST_VARIANTS(X) X(1, 5, 1) X(2, 4, 3) X(3, 4, 3)
