// Kernel variants: (coords per lane, active bodies per lane, all bodies per lane, sites per lane, joint slots per body).
// The dispatcher picks the first variant that fits the model; build.sh compiles one TU per entry.
#pragma once
#define STACB_VARIANTS(X) X(1, 1, 1, 1, 1) X(3, 1, 3, 1, 3) X(2, 2, 3, 1, 2) X(4, 3, 4, 2, 3) X(8, 6, 7, 2, 1)
