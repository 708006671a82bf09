// BN254 optimal-ate pairing, Fq12 helpers and the Groth16 verification equation — HOST code.
//
// In the reference this is CPU-only frontend code as well: `bn254_pairing`
// (/root/reference/icicle/src/pairing.cpp:11-25 -> include/icicle/pairing/models/bn.h:12-136,
// models/bls12.h:6-72,200-219), the `bn254_pairing_target_field_*` helpers
// (icicle/src/fields/ffi_extern_pairing_extension.cpp:6-50) and `groth16_verify_helper`
// (src/proof_helper.rs:319-372: 4 pairings on 4 threads).  It is not a fallback for anything on the GPU
// path; it exists so that the `verify` command and the Rust `pairing()` binding have a drop-in.
//
// What must match the reference bit for bit is the VALUE: e(P,Q) = f_{6x+2,Q}(P) * l_{[6x+2]Q,pi(Q)}(P) *
// l_{..,-pi^2(Q)}(P) raised to (p^12-1)/r * m, m = 2x(6x^2+3x+1) (the Fuentes-Castaneda et al. multiple the
// reference's hard part computes, bn.h:66-100), in the tower Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3-(9+u)),
// Fq12 = Fq6[w]/(w^2-v) (fields/snark_fields/bn254_tower.h:22-80), laid out c0.c0.c0 ... c1.c2.c1, standard form.
// How it is computed here differs: the Miller loop is fused (no coefficient vector), the loop digits are the
// NAF of 6x+2 derived at start-up, Frobenius constants are derived from xi^((p-1)/6) instead of tabulated,
// and the hard part is assembled from t^x, t^(x^2), t^(x^3) by the lambda_i of the paper.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "curve.cuh"
#include "host_math.h"

namespace b200 {
  namespace tower {

    static inline Fq2 mul_xi(const Fq2& t)
    {
      // (9 + u)(t0 + t1 u) = (9 t0 - t1) + (9 t1 + t0) u
      Fq2 t2 = t.dbl(), t4 = t2.dbl(), t8 = t4.dbl(), t9 = t8 + t;
      return {t9.c0 - t.c1, t9.c1 + t.c0};
    }
    static inline Fq2 scale(const Fq2& t, const Fq& k) { return {t.c0 * k, t.c1 * k}; }
    static inline Fq2 conj(const Fq2& t) { return {t.c0, t.c1.neg()}; }

    struct Fq6 { // a + b v + c v^2
      Fq2 a, b, c;
      static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
      static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
      bool is_zero() const { return a.is_zero() && b.is_zero() && c.is_zero(); }
      friend bool operator==(const Fq6& x, const Fq6& y) { return x.a == y.a && x.b == y.b && x.c == y.c; }
      friend Fq6 operator+(const Fq6& x, const Fq6& y) { return {x.a + y.a, x.b + y.b, x.c + y.c}; }
      friend Fq6 operator-(const Fq6& x, const Fq6& y) { return {x.a - y.a, x.b - y.b, x.c - y.c}; }
      Fq6 neg() const { return {a.neg(), b.neg(), c.neg()}; }
      Fq6 mul_v() const { return {mul_xi(c), a, b}; }
      Fq6 scaled(const Fq2& k) const { return {a * k, b * k, c * k}; }
      friend Fq6 operator*(const Fq6& x, const Fq6& y)
      {
        // Karatsuba over v, v^3 = xi
        Fq2 aa = x.a * y.a, bb = x.b * y.b, cc = x.c * y.c;
        Fq2 ab = (x.a + x.b) * (y.a + y.b) - aa - bb;
        Fq2 ac = (x.a + x.c) * (y.a + y.c) - aa - cc;
        Fq2 bc = (x.b + x.c) * (y.b + y.c) - bb - cc;
        return {aa + mul_xi(bc), ab + mul_xi(cc), ac + bb};
      }
      // x * (s + t v)
      Fq6 mul_sparse(const Fq2& s, const Fq2& t) const
      {
        Fq2 as = a * s, bt = b * t;
        Fq2 ab = (a + b) * (s + t) - as - bt; // a t + b s
        return {as + mul_xi(c * t), ab, c * s + bt};
      }
      Fq6 inverse() const
      {
        // adjugate of the multiplication matrix; norm to Fq2
        Fq2 A = a.sqr() - mul_xi(b * c);
        Fq2 B = mul_xi(c.sqr()) - a * b;
        Fq2 C = b.sqr() - a * c;
        Fq2 n = a * A + mul_xi(c * B + b * C);
        Fq2 ni = n.inverse();
        return {A * ni, B * ni, C * ni};
      }
    };

    struct Fq12 { // lo + hi w
      Fq6 lo, hi;
      static Fq12 zero() { return {Fq6::zero(), Fq6::zero()}; }
      static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
      friend bool operator==(const Fq12& x, const Fq12& y) { return x.lo == y.lo && x.hi == y.hi; }
      friend Fq12 operator+(const Fq12& x, const Fq12& y) { return {x.lo + y.lo, x.hi + y.hi}; }
      friend Fq12 operator-(const Fq12& x, const Fq12& y) { return {x.lo - y.lo, x.hi - y.hi}; }
      Fq12 conj() const { return {lo, hi.neg()}; } // the p^6 Frobenius
      friend Fq12 operator*(const Fq12& x, const Fq12& y)
      {
        Fq6 ll = x.lo * y.lo, hh = x.hi * y.hi;
        Fq6 cross = (x.lo + x.hi) * (y.lo + y.hi) - ll - hh;
        return {ll + hh.mul_v(), cross};
      }
      Fq12 sqr() const
      {
        // (lo + hi w)^2 = (lo + hi)(lo + v hi) - m - v m + 2 m w,  m = lo hi
        Fq6 m = lo * hi;
        Fq6 t = (lo + hi) * (lo + hi.mul_v()) - m - m.mul_v();
        return {t, m + m};
      }
      // this * (l0 + l3 w + l4 v w): a line through twist points evaluated at a G1 point
      Fq12 mul_line(const Fq2& l0, const Fq2& l3, const Fq2& l4) const
      {
        Fq6 t0 = lo.scaled(l0);
        Fq6 t1 = hi.mul_sparse(l3, l4);
        Fq6 t2 = (lo + hi).mul_sparse(l0 + l3, l4) - t0 - t1;
        return {t0 + t1.mul_v(), t2};
      }
      Fq12 inverse() const
      {
        Fq6 n = lo * lo - (hi * hi).mul_v();
        if (n.is_zero()) return zero();
        Fq6 ni = n.inverse();
        return {lo * ni, (hi * ni).neg()};
      }
    };

    // ---- constants derived once -----------------------------------------------------------------------------
    struct Consts {
      Fq2 gamma[4][6];  // gamma[k][i] = xi^(i (p^k - 1)/6): w^i -> gamma[k][i] w^i under the p^k Frobenius
      Fq2 b3_twist;     // 3 b' , b' = 3/xi
      Fq half;          // 1/2
      std::vector<int> loop_naf; // NAF of 6x+2, least significant first
      uint64_t x;       // BN parameter
    };

    static Fq2 fq2_pow(Fq2 base, const uint32_t* e, int words)
    {
      Fq2 acc = Fq2::one();
      for (int i = words - 1; i >= 0; --i) {
        for (int b = 31; b >= 0; --b) {
          acc = acc.sqr();
          if ((e[i] >> b) & 1) acc = acc * base;
        }
      }
      return acc;
    }

    static const Consts& consts()
    {
      static const Consts C = [] {
        Consts c;
        c.x = 0x44e992b44a6909f1ull; // 4965661367192848881 (pairing/params/bn254.h:13)
        // (p - 1) / 6 by schoolbook division of the limb array
        uint32_t e[8];
        uint64_t rem = 0;
        uint32_t pm1[8];
        for (int i = 0; i < 8; ++i)
          pm1[i] = FqCfg::P(i);
        pm1[0] -= 1; // p is odd
        for (int i = 7; i >= 0; --i) {
          uint64_t cur = (rem << 32) | pm1[i];
          e[i] = (uint32_t)(cur / 6);
          rem = cur % 6;
        }
        Fq nine = Fq::zero();
        nine.v[0] = 9;
        Fq2 xi = {Fq::to_mont(nine), Fq::one()};
        Fq2 g1 = fq2_pow(xi, e, 8);                // xi^((p-1)/6)
        Fq n = g1.c0.sqr() + g1.c1.sqr();          // g1 * conj(g1) = xi^((p^2-1)/6), in Fq
        Fq2 g2 = {n, Fq::zero()};
        Fq2 g3 = scale(g1, n);                     // xi^((p^3-1)/6)
        const Fq2 base[4] = {Fq2::one(), g1, g2, g3};
        for (int k = 0; k < 4; ++k) {
          c.gamma[k][0] = Fq2::one();
          for (int i = 1; i < 6; ++i)
            c.gamma[k][i] = c.gamma[k][i - 1] * base[k];
        }
        Fq three = Fq::zero();
        three.v[0] = 3;
        Fq2 b_twist = scale(xi.inverse(), Fq::to_mont(three));
        c.b3_twist = b_twist + b_twist + b_twist;
        Fq two = Fq::zero();
        two.v[0] = 2;
        c.half = Fq::to_mont(two).inverse();
        unsigned __int128 s = (unsigned __int128)c.x * 6 + 2;
        while (s) {
          int d = 0;
          if (s & 1) {
            d = ((s & 3) == 3) ? -1 : 1;
            if (d < 0)
              s += 1;
            else
              s -= 1;
          }
          c.loop_naf.push_back(d);
          s >>= 1;
        }
        return c;
      }();
      return C;
    }

    // Frobenius f -> f^(p^k), k = 1..3
    static Fq12 frobenius(const Fq12& f, int k)
    {
      const Consts& C = consts();
      auto map = [&](const Fq2& g, int i) { return ((k & 1) ? conj(g) : g) * C.gamma[k][i]; };
      // w-power of each slot: lo = w^0, w^2, w^4 ; hi = w^1, w^3, w^5
      return {{map(f.lo.a, 0), map(f.lo.b, 2), map(f.lo.c, 4)}, {map(f.hi.a, 1), map(f.hi.b, 3), map(f.hi.c, 5)}};
    }

    // ---- Miller loop ---------------------------------------------------------------------------------------
    struct TwistPoint { // homogeneous projective point on the D-type twist y^2 = x^3 + 3/xi
      Fq2 x, y, z;
    };

    // Costello-Lange-Naehrig doubling step; returns the tangent line (l0 to be scaled by yP, l3 by xP, l4)
    static void step_double(TwistPoint& r, Fq2& l0, Fq2& l3, Fq2& l4)
    {
      const Consts& C = consts();
      Fq2 a = scale(r.x * r.y, C.half);
      Fq2 b = r.y.sqr();
      Fq2 c = r.z.sqr();
      Fq2 e = C.b3_twist * c;
      Fq2 f = e + e + e;
      Fq2 g = scale(b + f, C.half);
      Fq2 h = (r.y + r.z).sqr() - (b + c);
      Fq2 j = r.x.sqr();
      Fq2 ee = e.sqr();
      l0 = h.neg();
      l3 = j + j + j;
      l4 = e - b;
      r.x = a * (b - f);
      r.y = g.sqr() - (ee + ee + ee);
      r.z = b * h;
    }

    // mixed addition step r += q; returns the chord line
    static void step_add(TwistPoint& r, const Fq2& qx, const Fq2& qy, Fq2& l0, Fq2& l3, Fq2& l4)
    {
      Fq2 theta = r.y - qy * r.z;
      Fq2 lambda = r.x - qx * r.z;
      Fq2 c = theta.sqr();
      Fq2 d = lambda.sqr();
      Fq2 e = lambda * d;
      Fq2 f = r.z * c;
      Fq2 g = r.x * d;
      Fq2 h = e + f - (g + g);
      l0 = lambda;
      l3 = theta.neg();
      l4 = theta * qx - lambda * qy;
      r.y = theta * (g - h) - e * r.y;
      r.x = lambda * h;
      r.z = r.z * e;
    }

    static Fq12 miller_loop(const G1Affine& p, const G2Affine& q) // Montgomery-form inputs
    {
      const Consts& C = consts();
      TwistPoint r = {q.x, q.y, Fq2::one()};
      Fq2 nqy = q.y.neg();
      Fq12 f = Fq12::one();
      Fq2 l0, l3, l4;
      auto absorb = [&] { f = f.mul_line(scale(l0, p.y), scale(l3, p.x), l4); };
      for (int i = (int)C.loop_naf.size() - 2; i >= 0; --i) {
        f = f.sqr();
        step_double(r, l0, l3, l4);
        absorb();
        int d = C.loop_naf[i];
        if (d) {
          step_add(r, q.x, d > 0 ? q.y : nqy, l0, l3, l4);
          absorb();
        }
      }
      // pi(Q) and -pi^2(Q) on the twist: (conj^k(x) gamma[k][2], conj^k(y) gamma[k][3])
      Fq2 q1x = conj(q.x) * C.gamma[1][2], q1y = conj(q.y) * C.gamma[1][3];
      Fq2 q2x = q.x * C.gamma[2][2], q2y = (q.y * C.gamma[2][3]).neg();
      step_add(r, q1x, q1y, l0, l3, l4);
      absorb();
      step_add(r, q2x, q2y, l0, l3, l4);
      absorb();
      return f;
    }

    // t^x for t in the cyclotomic subgroup (inverse = conjugate); plain MSB-first binary chain
    static Fq12 pow_x(const Fq12& t)
    {
      uint64_t x = consts().x;
      Fq12 acc = t;
      for (int b = 61; b >= 0; --b) { // x has 63 bits, top bit consumed by the initial value
        acc = acc.sqr();
        if ((x >> b) & 1) acc = acc * t;
      }
      return acc;
    }

    static Fq12 final_exponentiation(const Fq12& f)
    {
      Fq12 inv = f.inverse();
      if (inv == Fq12::zero()) return Fq12::one(); // the reference's answer for a non-invertible Miller value (bn.h:47-50)
      Fq12 t = f.conj() * inv;      // f^(p^6 - 1)
      t = frobenius(t, 2) * t;      // ^(p^2 + 1)
      // hard part: t^(l0 + l1 p + l2 p^2 + l3 p^3),
      //   l0 = 1 + 6x + 12x^2 + 12x^3, l1 = 4x + 6x^2 + 12x^3, l2 = 6x + 6x^2 + 12x^3, l3 = l1 - 1
      Fq12 a = pow_x(t), b = pow_x(a), c = pow_x(b);
      Fq12 a2 = a.sqr(), a4 = a2.sqr(), a6 = a4 * a2;
      Fq12 b2 = b.sqr(), b6 = b2.sqr() * b2;
      Fq12 c2 = c.sqr(), c4 = c2.sqr(), c12 = c4.sqr() * c4;
      Fq12 s = c12 * b6;
      Fq12 e1 = s * a4;
      Fq12 e2 = s * a6;
      Fq12 e3 = e1 * t.conj();
      Fq12 e0 = e2 * b6 * t;
      return e0 * frobenius(e1, 1) * frobenius(e2, 2) * frobenius(e3, 3);
    }

    // ---- boundary layout: 12 Fq in standard form, c0.c0.c0 .. c1.c2.c1 ---------------------------------------
    static Fq12 load_std(const uint32_t* in)
    {
      Fq2 g[6];
      for (int i = 0; i < 6; ++i) {
        memcpy(g[i].c0.v, in + 16 * i, 32);
        memcpy(g[i].c1.v, in + 16 * i + 8, 32);
        g[i] = Fq2::to_mont(g[i]);
      }
      return {{g[0], g[1], g[2]}, {g[3], g[4], g[5]}};
    }
    static void store_std(const Fq12& f, uint32_t* out)
    {
      const Fq2 g[6] = {f.lo.a, f.lo.b, f.lo.c, f.hi.a, f.hi.b, f.hi.c};
      for (int i = 0; i < 6; ++i) {
        Fq2 s = Fq2::from_mont(g[i]);
        memcpy(out + 16 * i, s.c0.v, 32);
        memcpy(out + 16 * i + 8, s.c1.v, 32);
      }
    }

    static Fq12 pairing_mont(const G1Affine& p, const G2Affine& q) { return final_exponentiation(miller_loop(p, q)); }

  } // namespace tower

  // ---- verification_key.json / proof.json / public.json (snarkjs layout; src/cache.rs:70-108, lib.rs:63-82) ---
  namespace vjson {
    struct Value {
      enum Kind { Null, Number, String, Array, Object } kind = Null;
      std::string text; // number or string payload
      std::vector<Value> items;
      std::vector<std::pair<std::string, Value>> fields;
      const Value* get(const char* key) const
      {
        for (auto& f : fields)
          if (f.first == key) return &f.second;
        return nullptr;
      }
    };

    struct Reader {
      const std::string& s;
      size_t i = 0;
      bool ok = true;
      explicit Reader(const std::string& str) : s(str) {}
      void ws()
      {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\r' || s[i] == '\t'))
          ++i;
      }
      bool eat(char c)
      {
        ws();
        if (i < s.size() && s[i] == c) {
          ++i;
          return true;
        }
        return false;
      }
      std::string str()
      {
        std::string out;
        if (!eat('"')) {
          ok = false;
          return out;
        }
        while (i < s.size() && s[i] != '"') {
          if (s[i] == '\\' && i + 1 < s.size()) ++i; // the files hold no escapes beyond \" and \\ at most
          out.push_back(s[i++]);
        }
        if (i >= s.size()) ok = false;
        ++i;
        return out;
      }
      Value value(int depth = 0)
      {
        Value v;
        ws();
        if (!ok || i >= s.size() || depth > 32) {
          ok = false;
          return v;
        }
        char c = s[i];
        if (c == '"') {
          v.kind = Value::String;
          v.text = str();
        } else if (c == '[') {
          ++i;
          v.kind = Value::Array;
          if (!eat(']')) {
            do {
              v.items.push_back(value(depth + 1));
            } while (ok && eat(','));
            if (!eat(']')) ok = false;
          }
        } else if (c == '{') {
          ++i;
          v.kind = Value::Object;
          if (!eat('}')) {
            do {
              std::string k = str();
              if (!eat(':')) ok = false;
              v.fields.emplace_back(k, value(depth + 1));
            } while (ok && eat(','));
            if (!eat('}')) ok = false;
          }
        } else if (c == '-' || (c >= '0' && c <= '9')) {
          v.kind = Value::Number;
          while (i < s.size() && (s[i] == '-' || s[i] == '+' || s[i] == '.' || s[i] == 'e' || s[i] == 'E' || (s[i] >= '0' && s[i] <= '9')))
            v.text.push_back(s[i++]);
        } else if (s.compare(i, 4, "true") == 0 || s.compare(i, 4, "null") == 0) {
          i += 4;
        } else if (s.compare(i, 5, "false") == 0) {
          i += 5;
        } else {
          ok = false;
        }
        return v;
      }
    };

    static bool read_file(const char* path, std::string& out)
    {
      std::ifstream f(path, std::ios::binary);
      if (!f) return false;
      std::ostringstream ss;
      ss << f.rdbuf();
      out = ss.str();
      return true;
    }

    // decimal string -> 8 little-endian 32-bit limbs (BigUint::parse_bytes(.., 10) + resize(32); conversions.rs:60-70)
    static bool decimal_to_limbs(const std::string& d, uint32_t out[8])
    {
      if (d.empty()) return false;
      uint32_t acc[8] = {0};
      for (char ch : d) {
        if (ch < '0' || ch > '9') return false;
        uint64_t carry = (uint64_t)(ch - '0');
        for (int i = 0; i < 8; ++i) {
          uint64_t cur = (uint64_t)acc[i] * 10 + carry;
          acc[i] = (uint32_t)cur;
          carry = cur >> 32;
        }
        if (carry) return false; // does not fit 256 bits
      }
      memcpy(out, acc, 32);
      return true;
    }

    // coordinates are JSON strings of decimal digits (serde: Vec<String>); a bare number is a type error there too
    static bool coord_of(const Value& v, uint32_t out[8]) { return v.kind == Value::String && decimal_to_limbs(v.text, out); }
    static bool g1_of(const Value* v, G1Affine& out_std)
    {
      if (!v || v->kind != Value::Array || v->items.size() < 2) return false;
      return coord_of(v->items[0], out_std.x.v) && coord_of(v->items[1], out_std.y.v);
    }
    static bool g2_of(const Value* v, G2Affine& out_std)
    {
      if (!v || v->kind != Value::Array || v->items.size() < 2) return false;
      const Value &x = v->items[0], &y = v->items[1];
      if (x.kind != Value::Array || y.kind != Value::Array || x.items.size() < 2 || y.items.size() < 2) return false;
      return coord_of(x.items[0], out_std.x.c0.v) && coord_of(x.items[1], out_std.x.c1.v) &&
             coord_of(y.items[0], out_std.y.c0.v) && coord_of(y.items[1], out_std.y.c1.v);
    }
  } // namespace vjson

  // e(-A, B) * e(cpub, gamma_2) * e(C, delta_2) * e(alpha_1, beta_2) == 1, cpub = IC_0 + sum pub_i IC_{i+1}
  // (src/proof_helper.rs:319-372).  All inputs in STANDARD form, as the JSON files hold them.
  // little-endian 8x32 value < modulus (P(k) = limb k)
  template <class P>
  static bool limbs_below(const uint32_t* v, P modulus)
  {
    for (int i = 7; i >= 0; --i) {
      uint32_t m = modulus(i);
      if (v[i] != m) return v[i] < m;
    }
    return false;
  }

  // y^2 == x^3 + b for a Montgomery-form affine point; (0,0) is the identity and passes
  template <class F>
  static bool affine_on_curve(const Affine<F>& p)
  {
    if (p.is_inf()) return true;
    return p.y.sqr() == p.x.sqr() * p.x + curve_b<F>();
  }

  static bool groth16_check(
    const G1Affine& a, const G2Affine& b, const G1Affine& c, const G1Affine& alpha1, const G2Affine& beta2, const G2Affine& gamma2,
    const G2Affine& delta2, const G1Affine* ic, const Fr* publics, size_t n_public)
  {
    XYZZ<Fq> cpub = XYZZ<Fq>::from_affine(affine_to_mont(ic[0]));
    for (size_t i = 0; i < n_public; ++i)
      cpub.add(host_scalar_mul(XYZZ<Fq>::from_affine(affine_to_mont(ic[i + 1])), publics[i]));
    G1Affine am = affine_to_mont(a);
    G1Affine neg_a = {am.x, am.y.neg()};
    const G1Affine ps[4] = {neg_a, cpub.to_affine(), affine_to_mont(c), affine_to_mont(alpha1)};
    const G2Affine qs[4] = {affine_to_mont(b), affine_to_mont(gamma2), affine_to_mont(delta2), affine_to_mont(beta2)};
    tower::consts(); // build the constants before the threads race for them
    tower::Fq12 e[4];
    std::thread th[4];
    for (int k = 0; k < 4; ++k)
      th[k] = std::thread([&, k] { e[k] = tower::pairing_mont(ps[k], qs[k]); });
    for (auto& t : th)
      t.join();
    return e[0] * e[1] * e[2] * e[3] == tower::Fq12::one();
  }

} // namespace b200

using namespace b200;

#define EXPORT __attribute__((visibility("default")))

extern "C" {

// icicle/src/pairing.cpp:20-24 (rust: icicle-core/src/pairing/mod.rs:37-44)
EXPORT void bn254_pairing(const bn254_affine_t* p, const bn254_g2_affine_t* q, bn254_fq12_t* out)
{
  G1Affine pm = affine_to_mont(*reinterpret_cast<const G1Affine*>(p));
  G2Affine qm = affine_to_mont(*reinterpret_cast<const G2Affine*>(q));
  tower::store_std(tower::pairing_mont(pm, qm), reinterpret_cast<uint32_t*>(out));
}

// icicle/src/fields/ffi_extern_pairing_extension.cpp:6-50
EXPORT void bn254_pairing_target_field_generate_scalars(bn254_fq12_t* out, int size)
{
  std::mt19937_64 rng(std::random_device{}());
  for (int n = 0; n < size; ++n) {
    uint32_t* w = reinterpret_cast<uint32_t*>(out + n);
    for (int k = 0; k < 12; ++k) {
      // uniform in [0, p) by rejection on 254 bits
      for (;;) {
        uint32_t v[8];
        for (int i = 0; i < 8; i += 2) {
          uint64_t r = rng();
          v[i] = (uint32_t)r;
          v[i + 1] = (uint32_t)(r >> 32);
        }
        v[7] &= 0x3fffffff;
        bool lt = false;
        for (int i = 7; i >= 0; --i) {
          if (v[i] != FqCfg::P(i)) {
            lt = v[i] < FqCfg::P(i);
            break;
          }
        }
        if (lt) {
          memcpy(w + 8 * k, v, 32);
          break;
        }
      }
    }
  }
}
EXPORT void bn254_pairing_target_field_add(const bn254_fq12_t* a, const bn254_fq12_t* b, bn254_fq12_t* out)
{
  tower::store_std(tower::load_std((const uint32_t*)a) + tower::load_std((const uint32_t*)b), (uint32_t*)out);
}
EXPORT void bn254_pairing_target_field_sub(bn254_fq12_t* a, bn254_fq12_t* b, bn254_fq12_t* out)
{
  tower::store_std(tower::load_std((const uint32_t*)a) - tower::load_std((const uint32_t*)b), (uint32_t*)out);
}
EXPORT void bn254_pairing_target_field_mul(const bn254_fq12_t* a, const bn254_fq12_t* b, bn254_fq12_t* out)
{
  tower::store_std(tower::load_std((const uint32_t*)a) * tower::load_std((const uint32_t*)b), (uint32_t*)out);
}
EXPORT void bn254_pairing_target_field_inv(const bn254_fq12_t* a, bn254_fq12_t* out)
{
  tower::store_std(tower::load_std((const uint32_t*)a).inverse(), (uint32_t*)out);
}
EXPORT void bn254_pairing_target_field_pow(const bn254_fq12_t* base, int exp, bn254_fq12_t* out)
{
  tower::Fq12 b = tower::load_std((const uint32_t*)base), acc = tower::Fq12::one();
  unsigned e = (unsigned)exp;
  while (e) {
    if (e & 1) acc = acc * b;
    b = b.sqr();
    e >>= 1;
  }
  tower::store_std(acc, (uint32_t*)out);
}
EXPORT void bn254_pairing_target_field_from_u32(uint32_t val, bn254_fq12_t* out)
{
  memset(out, 0, sizeof(*out));
  reinterpret_cast<uint32_t*>(out)[0] = val;
}

EXPORT eIcicleError b200_groth16_verify(
  const b200_groth16_proof* proof, const bn254_affine_t* vk_alpha_1, const bn254_g2_affine_t* vk_beta_2,
  const bn254_g2_affine_t* vk_gamma_2, const bn254_g2_affine_t* vk_delta_2, const bn254_affine_t* ic,
  const bn254_scalar_t* publics, uint64_t n_public, int* valid)
{
  if (!proof || !vk_alpha_1 || !vk_beta_2 || !vk_gamma_2 || !vk_delta_2 || !ic || !valid || (n_public && !publics))
    return ICICLE_INVALID_POINTER;
  // Input validation the reference's verifier (proof_helper.rs:319-372) does not do and snarkjs does: non-canonical
  // public inputs alias (x and x + r verify alike), non-canonical coordinates are proof malleability, and the pairing
  // is only sound for points on the curve / in the order-r subgroup of the twist.  Any violation: not valid.
  *valid = 0;
  for (uint64_t i = 0; i < n_public; ++i)
    if (!limbs_below(publics[i].limbs, [](int k) { return FrCfg::P(k); })) return ICICLE_SUCCESS;
  {
    const G1Affine& a = *reinterpret_cast<const G1Affine*>(&proof->pi_a);
    const G2Affine& b = *reinterpret_cast<const G2Affine*>(&proof->pi_b);
    const G1Affine& c = *reinterpret_cast<const G1Affine*>(&proof->pi_c);
    auto fq_ok = [](const Fq& x) { return limbs_below(x.v, [](int k) { return FqCfg::P(k); }); };
    if (!fq_ok(a.x) || !fq_ok(a.y) || !fq_ok(c.x) || !fq_ok(c.y) || !fq_ok(b.x.c0) || !fq_ok(b.x.c1) || !fq_ok(b.y.c0) ||
        !fq_ok(b.y.c1))
      return ICICLE_SUCCESS;
    if (!affine_on_curve(affine_to_mont(a)) || !affine_on_curve(affine_to_mont(c))) return ICICLE_SUCCESS;
    G2Affine bm = affine_to_mont(b);
    if (!affine_on_curve(bm)) return ICICLE_SUCCESS;
    // subgroup check on the twist: [r]B == O  (G1 has cofactor 1)
    Fr r_minus_1;
    for (int k = 0; k < 8; ++k)
      r_minus_1.v[k] = FrCfg::P(k);
    r_minus_1.v[0] -= 1; // r is odd
    XYZZ<Fq2> t = host_scalar_mul(XYZZ<Fq2>::from_affine(bm), r_minus_1);
    t.madd(bm);
    if (!t.is_inf()) return ICICLE_SUCCESS;
  }
  *valid = groth16_check(
             *reinterpret_cast<const G1Affine*>(&proof->pi_a), *reinterpret_cast<const G2Affine*>(&proof->pi_b),
             *reinterpret_cast<const G1Affine*>(&proof->pi_c), *reinterpret_cast<const G1Affine*>(vk_alpha_1),
             *reinterpret_cast<const G2Affine*>(vk_beta_2), *reinterpret_cast<const G2Affine*>(vk_gamma_2),
             *reinterpret_cast<const G2Affine*>(vk_delta_2), reinterpret_cast<const G1Affine*>(ic),
             reinterpret_cast<const Fr*>(publics), (size_t)n_public)
             ? 1
             : 0;
  return ICICLE_SUCCESS;
}

EXPORT eIcicleError b200_groth16_verify_files(const char* proof_path, const char* public_path, const char* vk_path, int* valid)
{
  using namespace vjson;
  if (!proof_path || !public_path || !vk_path || !valid) return ICICLE_INVALID_POINTER;
  *valid = 0;
  std::string text[3];
  const char* paths[3] = {proof_path, public_path, vk_path};
  Value doc[3];
  for (int k = 0; k < 3; ++k) {
    if (!read_file(paths[k], text[k])) return ICICLE_INVALID_ARGUMENT;
    Reader rd(text[k]);
    doc[k] = rd.value();
    rd.ws();
    if (!rd.ok || rd.i != text[k].size()) return ICICLE_INVALID_ARGUMENT; // malformed, or trailing characters
  }
  const Value &pj = doc[0], &pub = doc[1], &vk = doc[2];
  b200_groth16_proof proof;
  G1Affine alpha1;
  G2Affine beta2, gamma2, delta2;
  if (!g1_of(pj.get("pi_a"), *reinterpret_cast<G1Affine*>(&proof.pi_a)) ||
      !g2_of(pj.get("pi_b"), *reinterpret_cast<G2Affine*>(&proof.pi_b)) ||
      !g1_of(pj.get("pi_c"), *reinterpret_cast<G1Affine*>(&proof.pi_c)) || !g1_of(vk.get("vk_alpha_1"), alpha1) ||
      !g2_of(vk.get("vk_beta_2"), beta2) || !g2_of(vk.get("vk_gamma_2"), gamma2) || !g2_of(vk.get("vk_delta_2"), delta2))
    return ICICLE_INVALID_ARGUMENT;
  const Value* icv = vk.get("IC");
  const Value* npv = vk.get("nPublic");
  if (!icv || icv->kind != Value::Array || !npv || npv->kind != Value::Number || pub.kind != Value::Array) return ICICLE_INVALID_ARGUMENT;
  size_t n_public = (size_t)strtoull(npv->text.c_str(), nullptr, 10);
  // The reference takes `public.iter().take(n_public)` and zips it with IC, so a short public.json silently proves a
  // statement with the missing inputs read as 0 and a long one is truncated.  Here the three counts must agree.
  if (pub.items.size() != n_public || icv->items.size() != n_public + 1) return ICICLE_INVALID_ARGUMENT;
  size_t used = n_public;
  std::vector<G1Affine> ic(used + 1);
  for (size_t i = 0; i <= used; ++i)
    if (!g1_of(&icv->items[i], ic[i])) return ICICLE_INVALID_ARGUMENT;
  std::vector<Fr> publics(used);
  for (size_t i = 0; i < used; ++i)
    if (!coord_of(pub.items[i], publics[i].v)) return ICICLE_INVALID_ARGUMENT;
  return b200_groth16_verify(
    &proof, reinterpret_cast<const bn254_affine_t*>(&alpha1), reinterpret_cast<const bn254_g2_affine_t*>(&beta2),
    reinterpret_cast<const bn254_g2_affine_t*>(&gamma2), reinterpret_cast<const bn254_g2_affine_t*>(&delta2),
    reinterpret_cast<const bn254_affine_t*>(ic.data()), reinterpret_cast<const bn254_scalar_t*>(publics.data()), used, valid);
}

} // extern "C"
