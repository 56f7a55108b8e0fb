// tableaus.cuh -- the reference's explicit Runge-Kutta tableaus as compile-time data.
//
// Coefficients are the rationals of /root/reference src/runge_kutta.jl:66-157
// and become doubles the way the reference converts them on every solver call
// (Rational{Int64} -> Float64 = Float64(num)/Float64(den), src/ODE.jl:163,
// src/runge_kutta.jl:41-54).  The division happens in a constant expression, so
// each coefficient reaches SASS as an immediate / constant-bank operand and zero
// entries vanish from the unrolled stage loops at compile time.
//
// ORDER is minimum(btab.order) (src/runge_kutta.jl:252) -- the controller
// exponent is 1/(ORDER+1).  FSAL is isFSAL (src/runge_kutta.jl:62); it is true
// for dopri5 only (the reference's Bogacki-Shampine tableau steps with the
// 2nd-order row and is not FSAL).
#pragma once

namespace odeb200 {

struct Rat {
  long long n, d;
};
#define ODEB200_Z {0, 1}

template <class T>
struct TabOps {
  __host__ __device__ static constexpr double a(int i, int j) { return (double)T::A[i][j].n / (double)T::A[i][j].d; }
  __host__ __device__ static constexpr double b(int r, int j) { return (double)T::B[r][j].n / (double)T::B[r][j].d; }
  __host__ __device__ static constexpr double c(int i) { return (double)T::C[i].n / (double)T::C[i].d; }
  __host__ __device__ static constexpr bool fsal() {
    bool same = T::NB == 2;
    for (int j = 0; j < T::S; j++)
      if (a(T::S - 1, j) != b(0, j)) same = false;
    return same && c(T::S - 1) == 1.0;
  }
};

struct TabFEuler : TabOps<TabFEuler> {  // bt_feuler :66
  static constexpr int S = 1, NB = 1, ORDER = 1;
  static constexpr Rat A[1][1] = {{ODEB200_Z}};
  static constexpr Rat B[1][1] = {{{1, 1}}};
  static constexpr Rat C[1] = {ODEB200_Z};
};
struct TabMidpoint : TabOps<TabMidpoint> {  // bt_midpoint :71
  static constexpr int S = 2, NB = 1, ORDER = 2;
  static constexpr Rat A[2][2] = {{ODEB200_Z, ODEB200_Z}, {{1, 2}, ODEB200_Z}};
  static constexpr Rat B[1][2] = {{ODEB200_Z, {1, 1}}};
  static constexpr Rat C[2] = {ODEB200_Z, {1, 2}};
};
struct TabHeun : TabOps<TabHeun> {  // bt_heun :77
  static constexpr int S = 2, NB = 1, ORDER = 2;
  static constexpr Rat A[2][2] = {{ODEB200_Z, ODEB200_Z}, {{1, 1}, ODEB200_Z}};
  static constexpr Rat B[1][2] = {{{1, 2}, {1, 2}}};
  static constexpr Rat C[2] = {ODEB200_Z, {1, 1}};
};
struct TabRK4 : TabOps<TabRK4> {  // bt_rk4 :83
  static constexpr int S = 4, NB = 1, ORDER = 4;
  static constexpr Rat A[4][4] = {{ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
                                  {{1, 2}, ODEB200_Z, ODEB200_Z, ODEB200_Z},
                                  {ODEB200_Z, {1, 2}, ODEB200_Z, ODEB200_Z},
                                  {ODEB200_Z, ODEB200_Z, {1, 1}, ODEB200_Z}};
  static constexpr Rat B[1][4] = {{{1, 6}, {1, 3}, {1, 3}, {1, 6}}};
  static constexpr Rat C[4] = {ODEB200_Z, {1, 2}, {1, 2}, {1, 1}};
};
struct TabRK21 : TabOps<TabRK21> {  // bt_rk21 :93, order (2,1)
  static constexpr int S = 2, NB = 2, ORDER = 1;
  static constexpr Rat A[2][2] = {{ODEB200_Z, ODEB200_Z}, {{1, 1}, ODEB200_Z}};
  static constexpr Rat B[2][2] = {{{1, 2}, {1, 2}}, {{1, 1}, ODEB200_Z}};
  static constexpr Rat C[2] = {ODEB200_Z, {1, 1}};
};
struct TabRK23 : TabOps<TabRK23> {  // bt_rk23 :101, order (2,3)
  static constexpr int S = 4, NB = 2, ORDER = 2;
  static constexpr Rat A[4][4] = {{ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
                                  {{1, 2}, ODEB200_Z, ODEB200_Z, ODEB200_Z},
                                  {ODEB200_Z, {3, 4}, ODEB200_Z, ODEB200_Z},
                                  {{2, 9}, {1, 3}, {4, 9}, ODEB200_Z}};
  static constexpr Rat B[2][4] = {{{7, 24}, {1, 4}, {1, 3}, {1, 8}}, {{2, 9}, {1, 3}, {4, 9}, ODEB200_Z}};
  static constexpr Rat C[4] = {ODEB200_Z, {1, 2}, {3, 4}, {1, 1}};
};
struct TabRK45 : TabOps<TabRK45> {  // bt_rk45 (Fehlberg) :112, order (4,5)
  static constexpr int S = 6, NB = 2, ORDER = 4;
  static constexpr Rat A[6][6] = {
      {ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{1, 4}, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{3, 32}, {9, 32}, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{1932, 2197}, {-7200, 2197}, {7296, 2197}, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{439, 216}, {-8, 1}, {3680, 513}, {-845, 4104}, ODEB200_Z, ODEB200_Z},
      {{-8, 27}, {2, 1}, {-3544, 2565}, {1859, 4104}, {-11, 40}, ODEB200_Z}};
  static constexpr Rat B[2][6] = {{{25, 216}, ODEB200_Z, {1408, 2565}, {2197, 4104}, {-1, 5}, ODEB200_Z},
                                  {{16, 135}, ODEB200_Z, {6656, 12825}, {28561, 56430}, {-9, 50}, {2, 55}}};
  static constexpr Rat C[6] = {ODEB200_Z, {1, 4}, {3, 8}, {12, 13}, {1, 1}, {1, 2}};
};
// Cash-Karp 4(5).  The reference's README (README.md:57) says "Fehlberg and Cash-Karp coefficients are also available"
// for ode45, but src/runge_kutta.jl defines no such tableau: this one is an EXTENSION with the
// coefficients of Cash & Karp, ACM TOMS 16 (1990) 201-222, laid out like bt_dopri5 -- order (5,4), the solver advances
// with the 5th-order row b[1,:] and the controller exponent is 1/(min(order)+1) = 1/5.  Not FSAL (c[end] = 7/8).
struct TabCK45 : TabOps<TabCK45> {
  static constexpr int S = 6, NB = 2, ORDER = 4;
  static constexpr Rat A[6][6] = {
      {ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{1, 5}, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{3, 40}, {9, 40}, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{3, 10}, {-9, 10}, {6, 5}, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{-11, 54}, {5, 2}, {-70, 27}, {35, 27}, ODEB200_Z, ODEB200_Z},
      {{1631, 55296}, {175, 512}, {575, 13824}, {44275, 110592}, {253, 4096}, ODEB200_Z}};
  static constexpr Rat B[2][6] = {{{37, 378}, ODEB200_Z, {250, 621}, {125, 594}, ODEB200_Z, {512, 1771}},
                                  {{2825, 27648}, ODEB200_Z, {18575, 48384}, {13525, 55296}, {277, 14336}, {1, 4}}};
  static constexpr Rat C[6] = {ODEB200_Z, {1, 5}, {3, 10}, {3, 5}, {1, 1}, {7, 8}};
};
struct TabDopri5 : TabOps<TabDopri5> {  // bt_dopri5 :124, order (5,4)
  static constexpr int S = 7, NB = 2, ORDER = 4;
  static constexpr Rat A[7][7] = {
      {ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{1, 5}, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{3, 40}, {9, 40}, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{44, 45}, {-56, 15}, {32, 9}, ODEB200_Z, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{19372, 6561}, {-25360, 2187}, {64448, 6561}, {-212, 729}, ODEB200_Z, ODEB200_Z, ODEB200_Z},
      {{9017, 3168}, {-355, 33}, {46732, 5247}, {49, 176}, {-5103, 18656}, ODEB200_Z, ODEB200_Z},
      {{35, 384}, ODEB200_Z, {500, 1113}, {125, 192}, {-2187, 6784}, {11, 84}, ODEB200_Z}};
  static constexpr Rat B[2][7] = {
      {{35, 384}, ODEB200_Z, {500, 1113}, {125, 192}, {-2187, 6784}, {11, 84}, ODEB200_Z},
      {{5179, 57600}, ODEB200_Z, {7571, 16695}, {393, 640}, {-92097, 339200}, {187, 2100}, {1, 40}}};
  static constexpr Rat C[7] = {ODEB200_Z, {1, 5}, {3, 10}, {4, 5}, {8, 9}, {1, 1}, {1, 1}};
};
struct TabFeh78 : TabOps<TabFeh78> {  // bt_feh78 :140, order (7,8)
  static constexpr int S = 13, NB = 2, ORDER = 7;
#define Z ODEB200_Z
  static constexpr Rat A[13][13] = {
      {Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z},
      {{2, 27}, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z},
      {{1, 36}, {1, 12}, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z},
      {{1, 24}, Z, {1, 8}, Z, Z, Z, Z, Z, Z, Z, Z, Z, Z},
      {{5, 12}, Z, {-25, 16}, {25, 16}, Z, Z, Z, Z, Z, Z, Z, Z, Z},
      {{1, 20}, Z, Z, {1, 4}, {1, 5}, Z, Z, Z, Z, Z, Z, Z, Z},
      {{-25, 108}, Z, Z, {125, 108}, {-65, 27}, {125, 54}, Z, Z, Z, Z, Z, Z, Z},
      {{31, 300}, Z, Z, Z, {61, 225}, {-2, 9}, {13, 900}, Z, Z, Z, Z, Z, Z},
      {{2, 1}, Z, Z, {-53, 6}, {704, 45}, {-107, 9}, {67, 90}, {3, 1}, Z, Z, Z, Z, Z},
      {{-91, 108}, Z, Z, {23, 108}, {-976, 135}, {311, 54}, {-19, 60}, {17, 6}, {-1, 12}, Z, Z, Z, Z},
      {{2383, 4100}, Z, Z, {-341, 164}, {4496, 1025}, {-301, 82}, {2133, 4100}, {45, 82}, {45, 164}, {18, 41}, Z, Z, Z},
      {{3, 205}, Z, Z, Z, Z, {-6, 41}, {-3, 205}, {-3, 41}, {3, 41}, {6, 41}, Z, Z, Z},
      {{-1777, 4100}, Z, Z, {-341, 164}, {4496, 1025}, {-289, 82}, {2193, 4100}, {51, 82}, {33, 164}, {12, 41}, Z, {1, 1}, Z}};
  static constexpr Rat B[2][13] = {
      {{41, 840}, Z, Z, Z, Z, {34, 105}, {9, 35}, {9, 35}, {9, 280}, {9, 280}, {41, 840}, Z, Z},
      {Z, Z, Z, Z, Z, {34, 105}, {9, 35}, {9, 35}, {9, 280}, {9, 280}, Z, {41, 840}, {41, 840}}};
  static constexpr Rat C[13] = {Z, {2, 27}, {1, 9}, {1, 6}, {5, 12}, {1, 2}, {5, 6}, {1, 6}, {2, 3}, {1, 3}, {1, 1}, Z, {1, 1}};
#undef Z
};

}  // namespace odeb200
