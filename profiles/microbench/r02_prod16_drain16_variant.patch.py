p='/root/repo/openfermion-fqe_b200/csrc/ozaki.cu'
s=open(p).read()

old="""// predicated 16-byte load of four source-table entries, OZ_NONE when off"""
new="""// predicated 16-byte load of the (re, im) digit words of one determinant, `dflt` twice when off
__device__ __forceinline__ ulonglong2 oz_ldg128q_if(bool pred, const ulonglong2 *p, uint64_t dflt) {
  ulonglong2 v = make_ulonglong2(dflt, dflt);
  asm volatile("{\\n .reg .pred q;\\n setp.ne.b32 q, %3, 0;\\n @q ld.global.nc.v2.u64 {%0,%1}, [%2];\\n}"
               : "+l"(v.x), "+l"(v.y)
               : "l"(p), "r"((int)pred));
  return v;
}
#define OZ_TMEM_LD16x256(r, addr)                                                  \\
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"      \\
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])                   \\
               : "r"(addr))
__device__ __forceinline__ void oz_stcs_c128(double2 *ptr, double re, double im) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(ptr), "d"(re), "d"(im) : "memory");
}
// predicated 16-byte load of four source-table entries, OZ_NONE when off"""
assert old in s; s=s.replace(old,new,1)

old="""// KC: number of 16-byte K columns of the pair space (9 at norb = 16, 7 at norb = 14 with"""
new=r'''#ifdef OZ_PROD16
// Variant (OZ_PROD16): tile row of (ar, bc, part) = 16 ar + 8 part + bc.
// Producer: one thread per DETERMINANT and K sixth, 16-byte loads that bring the real and the
// imaginary digit word of a source at once, work unit = a quad of 4 pair indices.
template <bool PROF, int KC>
__device__ __forceinline__ void oz_produce16(const OzParams &p, uint8_t *s_d, uint32_t bar_full,
                                             uint32_t bar_free, int64_t my_tiles, int tid) {
  static_assert(KC > 0 && (2 * KC) % 6 == 0, "quads must divide evenly over six thread groups");
  constexpr int NQ = (4 * KC) / 6;   // quads per thread (6 at KC = 9)
  const int det = tid & 63, h = tid >> 6;
  const int bc = det & 7, ar = det >> 3;
  const ulonglong2 *pl = reinterpret_cast<const ulonglong2 *>(p.planes);
  const ulonglong2 *plT = reinterpret_cast<const ulonglong2 *>(p.planesT);
  const uint64_t ZERO = 0x4040404040404040ull, K128 = 0x8080808080808080ull;
  const uint32_t dst_re = oz_smem_u32(s_d) + (uint32_t)(2 * ar) * 128 + (uint32_t)bc * 16;
  long long c_wait = 0, c_prod = 0;
  for (int64_t it = 0; it < my_tiles; ++it) {
    const int64_t tile = blockIdx.x + it * gridDim.x;
    const int r = (int)(tile / p.tiles_per_row);
    const int bt = (int)(tile - (int64_t)r * p.tiles_per_row);
    const int64_t a_loc = 8 * (int64_t)r + ar, b = 8 * (int64_t)bt + bc;
    const bool valid = a_loc < p.nrows && b < p.lenb;
    const uint32_t a = (uint32_t)(p.row0 + (valid ? a_loc : 0)), bb = valid ? (uint32_t)b : 0u;
    const uint32_t *ta_row = p.srcT_a + oz_table_offset(a, 0, p.kpad);
    const uint32_t *tb_row = p.srcT_b + oz_table_offset(bb, 0, p.kpad);
    auto load_srcs = [&](int qq, uint32_t (&ta_)[4], uint32_t (&tb_)[4]) {
      const int ofs = 32 * (qq >> 1) + 4 * (qq & 1);
      oz_ldg128_if(valid, ta_row + ofs, ta_[0], ta_[1], ta_[2], ta_[3]);
      oz_ldg128_if(valid, tb_row + ofs, tb_[0], tb_[1], tb_[2], tb_[3]);
    };
    auto load_digits = [&](const uint32_t (&ta_)[4], const uint32_t (&tb_)[4], ulonglong2 (&va)[4],
                           ulonglong2 (&vb)[4]) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        va[u] = oz_ldg128q_if(ta_[u] != OZ_NONE, pl + (uint64_t)(ta_[u] + bb), ZERO);
        vb[u] = oz_ldg128q_if(tb_[u] != OZ_NONE, plT + (uint64_t)(tb_[u] + a), ZERO);
      }
    };
    auto transpose_store = [&](const uint64_t (&e)[4], uint32_t dst) {
      const uint32_t l0 = (uint32_t)e[0], l1 = (uint32_t)e[1], l2 = (uint32_t)e[2], l3 = (uint32_t)e[3];
      const uint32_t h0 = (uint32_t)(e[0] >> 32), h1 = (uint32_t)(e[1] >> 32),
                     h2 = (uint32_t)(e[2] >> 32), h3 = (uint32_t)(e[3] >> 32);
      const uint32_t t0 = __byte_perm(l0, l1, 0x5140), t1 = __byte_perm(l2, l3, 0x5140);
      const uint32_t t2 = __byte_perm(l0, l1, 0x7362), t3 = __byte_perm(l2, l3, 0x7362);
      const uint32_t t4 = __byte_perm(h0, h1, 0x5140), t5 = __byte_perm(h2, h3, 0x5140);
      uint32_t out[OZ_NS];
      out[0] = __byte_perm(t0, t1, 0x5410);
      out[1] = __byte_perm(t0, t1, 0x7632);
      out[2] = __byte_perm(t2, t3, 0x5410);
      out[3] = __byte_perm(t2, t3, 0x7632);
      out[4] = __byte_perm(t4, t5, 0x5410);
      out[5] = __byte_perm(t4, t5, 0x7632);
#pragma unroll
      for (int sl = 0; sl < OZ_NS; ++sl)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + (uint32_t)(sl * KC) * 2048), "r"(out[sl])
                     : "memory");
    };
    auto combine_store = [&](const ulonglong2 (&va)[4], const ulonglong2 (&vb)[4], int qq) {
      const uint32_t dst = dst_re + (uint32_t)(qq >> 2) * 2048 + 4u * (uint32_t)(qq & 3);
      uint64_t e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) e[u] = (va[u].x + vb[u].x) ^ K128;
      transpose_store(e, dst);
#pragma unroll
      for (int u = 0; u < 4; ++u) e[u] = (va[u].y + vb[u].y) ^ K128;
      transpose_store(e, dst + 128);   // imaginary part: row 16 ar + 8 + bc
    };
    const long long c_p0 = oz_clock<PROF>();
    uint32_t sa[2][4], sb[2][4];
    ulonglong2 va[4], vb[4];
    load_srcs(NQ * h, sa[0], sb[0]);
    load_srcs(NQ * h + 1, sa[1], sb[1]);
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      load_digits(sa[j & 1], sb[j & 1], va, vb);
      if (j + 2 < NQ) load_srcs(NQ * h + j + 2, sa[j & 1], sb[j & 1]);
      if (j == 0) {
        const long long c_w = oz_clock<PROF>();
        if (it >= 1) oz_mbar_wait(bar_free, (uint32_t)((it - 1) & 1));
        c_wait += oz_clock<PROF>() - c_w;
      }
      combine_store(va, vb, NQ * h + j);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    oz_mbar_arrive(bar_full);
    c_prod += oz_clock<PROF>() - c_p0;
  }
  if (PROF && p.prof && tid == 0) {
    atomicAdd(p.prof + 3, (unsigned long long)c_wait);
    atomicAdd(p.prof + 4, (unsigned long long)(c_prod - c_wait));
  }
}

// Drain of the same variant: tcgen05.ld.16x256b hands a thread rows g and g + 8 of a 16-row group
// (= the real and the imaginary part of determinant (ar, bc = g)) for two adjacent columns, so E
// leaves in 16-byte stores: half the store instructions of the row-per-lane drain.
template <bool PROF>
__device__ __forceinline__ void oz_drain16(const OzParams &p, uint32_t tmem, uint32_t bar_sfull,
                                           uint32_t bar_sfree, int64_t my_tiles, int nblk, int warp,
                                           int lane, int tid) {
  const int q = warp & 3;                       // TMEM lane quarter of this warp
  const int ch = (warp - OZ2_W_DRAIN) >> 2;     // column half of a block
  const int g = lane >> 2, t4 = lane & 3;
  const double st = p.stats[2] * p.op_scale;
  double w[3];
  {
    const double r = 1.0 / (double)OZ_RADIX;
    w[0] = st * r * r * r;
    w[1] = w[0] * r * r;
    w[2] = w[1] * r * r;
  }
  long long c_drain = 0, c_store = 0;
  for (int64_t it = 0; it < my_tiles; ++it) {
    const int64_t tile = blockIdx.x + it * gridDim.x;
    const int r = (int)(tile / p.tiles_per_row);
    const int bt = (int)(tile - (int64_t)r * p.tiles_per_row);
    const int64_t b = 8 * (int64_t)bt + g;
#pragma unroll 1
    for (int nb = 0; nb < nblk; ++nb) {
      constexpr int NCOL = OZ2_NB / 2;   // 24 columns per warp: 3 groups of 8
      const long long c_d0 = oz_clock<PROF>();
      const uint32_t par = (uint32_t)((nblk * it + nb) & 1);
      double run[2][NCOL / 8][4];        // [16-row half][column group][(col & 1) + 2 part]
#pragma unroll
      for (int pr = 0; pr < 3; ++pr) {
        const int sa = 2 * pr, sb = 2 * pr + 1;
        oz_mbar_wait(bar_sfull + 8 * sa, par);
        oz_mbar_wait(bar_sfull + 8 * sb, par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
          for (int rd = 0; rd < NCOL / 8; ++rd) {
            uint32_t ra[4], rb[4];
            const uint32_t col = tmem + ((uint32_t)(q * 32 + 16 * hf) << 16) +
                                 (uint32_t)(OZ2_A_COLS + NCOL * ch + 8 * rd);
            OZ_TMEM_LD16x256(ra, col + (uint32_t)(sa * OZ2_NB));
            OZ_TMEM_LD16x256(rb, col + (uint32_t)(sb * OZ2_NB));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int comb = (int)ra[v] * OZ_RADIX + (int)rb[v];
              const double cd = __hiloint2double(0x43300000, comb ^ (int)0x80000000) -
                                4503601774854144.0;
              run[hf][rd][v] = pr == 0 ? w[0] * cd : fma(w[pr], cd, run[hf][rd][v]);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          oz_mbar_arrive(bar_sfree + 8 * sa);
          oz_mbar_arrive(bar_sfree + 8 * sb);
        }
      }
      const long long c_d1 = oz_clock<PROF>();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int64_t a_loc = 8 * (int64_t)r + 2 * q + hf;
        if (a_loc < p.nrows && b < p.lenb) {
#pragma unroll
          for (int rd = 0; rd < NCOL / 8; ++rd)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int kl = OZ2_NB * nb + NCOL * ch + 8 * rd + 2 * t4 + c;
              if (kl < p.np)
                oz_stcs_c128(p.E + ((int64_t)kl * p.lde + a_loc * p.pitch + b), run[hf][rd][c],
                             run[hf][rd][c + 2]);
            }
        }
      }
      c_drain += c_d1 - c_d0;
      c_store += oz_clock<PROF>() - c_d1;
    }
  }
  if (PROF && p.prof && tid == OZ2_PRODUCERS) {
    atomicAdd(p.prof + 5, (unsigned long long)c_drain);
    atomicAdd(p.prof + 6, (unsigned long long)c_store);
  }
}
#endif

// KC: number of 16-byte K columns of the pair space (9 at norb = 16, 7 at norb = 14 with'''
assert old in s; s=s.replace(old,new,1)

old="""    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;\\n");
    
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch (warp % 4)"""
new="""    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;\\n");
#ifdef OZ_PROD16
    if constexpr (KC > 0 && (2 * KC) % 6 == 0) {
      oz_drain16<PROF>(p, tmem, bar_sfull, bar_sfree, my_tiles, nblk, warp, lane, tid);
    } else
#endif
    {
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch (warp % 4)"""
assert old in s; s=s.replace(old,new,1)
old="""    if (PROF && p.prof && tid == OZ2_PRODUCERS) {
      atomicAdd(p.prof + 5, (unsigned long long)c_drain);
      atomicAdd(p.prof + 6, (unsigned long long)c_store);
    }
  } else {
    // ================================= producers ======================================="""
new="""    if (PROF && p.prof && tid == OZ2_PRODUCERS) {
      atomicAdd(p.prof + 5, (unsigned long long)c_drain);
      atomicAdd(p.prof + 6, (unsigned long long)c_store);
    }
    }
  } else {
    // ================================= producers ======================================="""
assert old in s; s=s.replace(old,new,1)

old="""    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;\\n");
    // A tile is an 8 x 8 block of determinants (8 alpha rows x 8 beta columns) x (re, im).  The 32"""
new="""    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;\\n");
#ifdef OZ_PROD16
    if constexpr (KC > 0 && (2 * KC) % 6 == 0) {
      oz_produce16<PROF, KC>(p, s_d, bar_full, bar_free, my_tiles, tid);
    } else
#endif
    {
    // A tile is an 8 x 8 block of determinants (8 alpha rows x 8 beta columns) x (re, im).  The 32"""
assert old in s; s=s.replace(old,new,1)
old="""    if (PROF && p.prof && tid == 0) {
      atomicAdd(p.prof + 3, (unsigned long long)c_wait);
      atomicAdd(p.prof + 4, (unsigned long long)(c_prod - c_wait));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == OZ2_W_ISSUE)"""
new="""    if (PROF && p.prof && tid == 0) {
      atomicAdd(p.prof + 3, (unsigned long long)c_wait);
      atomicAdd(p.prof + 4, (unsigned long long)(c_prod - c_wait));
    }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == OZ2_W_ISSUE)"""
assert old in s; s=s.replace(old,new,1)
open(p,'w').write(s)
print("patched")
