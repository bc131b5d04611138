// ps_types.cuh -- device-visible constants and pore-model records shared by the kernel files.
#pragma once

namespace psdev {

constexpr int    N_STATES = 1024;
constexpr double NEG      = -1e300;       // cpp/AlignUtil.h:20 ("inf" is 1e300)

// packed step byte written by the forward fill: low 3 bits main-matrix move, bits 3-4 stay-matrix move
// main : 0 skip, 1 match, 2 insert, 3 ignore, 4 stay, 6 implicit (cpp/Alignment.cpp:19-28), 7 = score <= 0
// stay : 0 = score <= 0, 1 = stay, 2 = extend
constexpr int ST_SKIP = 0, ST_MATCH = 1, ST_INSERT = 2, ST_IGNORE = 3, ST_STAY = 4, ST_IMPLICIT = 6, ST_STOP = 7;

struct StateParams            // one 64-byte record per 5-mer state (cpp/EventData.h:21-45)
{
    double lev_mean, lev_stdv, log_lev, sd_mean, sd_lambda, log_lambda;
    double r_lev_stdv, r_sd_mean;      // correctly rounded reciprocals of the two divisors (host IEEE division)
};

struct __align__(16) LevelRec // one 32-byte record per event level (cpp/EventData.h:96-101, :218-220)
{
    double mean, stdv;
    double rstdv;                      // RN(1 / stdv)
    double lsd3;                       // 3 * log(stdv)
};

struct LevIn                  // what the host stages per level: 24 bytes; k_rows expands it into the records above/below
{
    double mean, stdv;
    double lsd3;                       // 3 * log(stdv), host libm (the only transcendental of the path, cpp/EventData.h:220)
};

struct StateParamsF           // FP32 fused emission coefficients of one state (ps_fast.cuh)
{
    float mu;                          // lev_mean
    float a_s;                         // -0.5 / lev_stdv^2
    float c_s;                         // -0.5 log2pi - log lev_stdv + 0.5 (log lambda - log2pi) + lik_offset
    float mu2;                         // sd_mean
    float f_s;                         // -0.5 lambda / sd_mean^2
    float pad0, pad1, pad2;
};

struct __align__(16) LevelRecF // FP32 row record of the forward scan, row i (1-based) at index i-1: mean, stdv, 1/stdv of
{                              // level i-1 and -1.5 log(stdv) of level n0-i (quirk A.3-1: cpp/Alignment.cpp:171-172)
    float x, y, ry, ey;
};

struct RegTabDev              // per region: first mutation, its events
{
    long long mut_off; int ev0, nev;
    int plain_points, pad;             // the region's mutations are the implicit point edits of an all-ACGT sequence: k_points writes them
};

struct ModelDev               // cpp/EventData.h:21-74
{
    StateParams st[N_STATES];
    double lskip, lstay, lext, lins;
};


} // namespace psdev
