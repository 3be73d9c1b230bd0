// fj_kernels.cuh — the sm_100a kernels of the B200-native Fujiyama hot path.
//
//   k_render_samples   persistent-threads sample kernel: sampler -> camera ray -> closest hit ->
//                      device shaders -> secondary/shadow rays (per-path DFS), one camera sample per lane,
//                      warps pull 8x4 sample blocks from a global work counter
//                      (render_tile/integrate_samples, src/fj_renderer.cc:1061-1121; SlTrace, src/fj_shading.cc:140)
//   k_resolve_tiles    Gaussian pixel filter of one tile's samples into packed tile blocks
//                      (reconstruct_image/apply_pixel_filter, src/fj_renderer.cc:939-995)
//   k_trace_closest    probe: closest hit of caller-supplied rays (Accelerator::Intersect)
//
// All device math lives in fj_device.cuh.  This translation unit is compiled with -fmad=false: every FMA
// in it is an explicit fmaf()/fma() call.
#pragma once

#include "fj_device.cuh"

namespace fj {

// ------------------------------------------------------------------------------------------ streaming access
// Ray / hit records and accumulators are touched once per round and are far larger than L2; the BVH is re-read by every
// ray and almost fits in L2.  Stream the former with evict-first hints so they do not push the latter out.
__device__ __forceinline__ void store_ray_cs(RayRec *dst, const RayRec &r) {
  const uint4 *s = reinterpret_cast<const uint4 *>(&r); uint4 *d = reinterpret_cast<uint4 *>(dst);
#pragma unroll
  for (int k = 0; k < 7; k++) __stcs(d + k, s[k]);
}
__device__ __forceinline__ void load_ray_cs(RayRec *dst, const RayRec *src) {
  uint4 *d = reinterpret_cast<uint4 *>(dst); const uint4 *s = reinterpret_cast<const uint4 *>(src);
#pragma unroll
  for (int k = 0; k < 7; k++) d[k] = __ldcs(s + k);
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void store_hit_cs(HitRec *dst, const HitRec &h) {
  const uint4 *s = reinterpret_cast<const uint4 *>(&h); uint4 *d = reinterpret_cast<uint4 *>(dst);
  __stcs(d, s[0]); __stcs(d + 1, s[1]);
}
__device__ __forceinline__ void load_hit_cs(HitRec *dst, const HitRec *src) {
  uint4 *d = reinterpret_cast<uint4 *>(dst); const uint4 *s = reinterpret_cast<const uint4 *>(src);
  d[0] = __ldcs(s); d[1] = __ldcs(s + 1);
}

// ------------------------------------------------------------------------------------------ sampler
// Tile sample grid, FixedGridSampler::generate_samples (src/fj_fixed_grid_sampler.cc:33-84).
struct TileGrid { int nsx, nsy, xoff, yoff, nbx, nby; };
__device__ __forceinline__ TileGrid tile_grid(const DFrame &fr, const DTile &t) {
  TileGrid g;
  g.nsx = fr.xrate * (t.xmax - t.xmin) + 2 * fr.mx;
  g.nsy = fr.yrate * (t.ymax - t.ymin) + 2 * fr.my;
  g.xoff = t.xmin * fr.xrate - fr.mx;
  g.yoff = t.ymin * fr.yrate - fr.my;
  g.nbx = (g.nsx + 7) >> 3;
  g.nby = (g.nsy + 3) >> 2;
  return g;
}
// Slot of sample (x, y) inside the tile's sample buffer: 8x4 sample blocks, one block per warp fetch.
__device__ __forceinline__ uint32_t sample_slot(const TileGrid &g, int x, int y) {
  return (uint32_t)((((y >> 2) * g.nbx + (x >> 3)) << 5) + ((y & 3) << 3) + (x & 7));
}
__device__ __forceinline__ void sample_uv(const DFrame &fr, const TileGrid &g, int x, int y, double *u, double *v) {
  // s->u = (.5 + x + xoffset) * udelta ; s->v = 1 - (.5 + y + yoffset) * vdelta        (:63-64)
  double su = dmul(dadd(dadd(.5, (double)x), (double)g.xoff), fr.udelta);
  double sv = dsub(1., dmul(dadd(dadd(.5, (double)y), (double)g.yoff), fr.vdelta));
  if (fr.jitter > 0) {                                                               // (:66-72)
    const uint32_t k = (uint32_t)(y * g.nsx + x);
    const double uj = dmul(ddiv((double)fr.jitter_tab[2 * k], 4294967295.0), fr.jitter);
    const double vj = dmul(ddiv((double)fr.jitter_tab[2 * k + 1], 4294967295.0), fr.jitter);
    su = dadd(su, dmul(fr.udelta, dsub(uj, .5)));
    sv = dadd(sv, dmul(fr.vdelta, dsub(vj, .5)));
  }
  *u = su; *v = sv;
}
// Camera::GetRay, src/fj_camera.cc:79-110; `tidx` = the sample's entry of the time table (its matrix when the camera moves)
__device__ __forceinline__ void camera_ray(const DCamera &c, double u, double v, uint32_t tidx, RayD *ray) {
  const double *fwd = c.motion ? c.motion + 12 * (size_t)tidx : c.fwd;
  const D3 target = mat_point(fwd, mk(dmul(dsub(u, .5), c.uvx), dmul(dsub(v, .5), c.uvy), -1.));
  const D3 eye = mat_point(fwd, mk(0., 0., 0.));
  ray->d = normalize(target - eye); ray->o = eye; ray->tmin = c.znear; ray->tmax = c.zfar; ray->tidx = tidx;
}

// ------------------------------------------------------------------------------------------ lights
struct LightSample { D3 P, N; C3 color; };

// Light::Illuminate of the four light types (fj_point_light.cc:41-44, fj_rectangle_light.cc:49-61,
// fj_sphere_light.cc:48-60, fj_dome_light.cc:54-57).
__device__ __forceinline__ C3 light_illuminate(const DLight &lt, const LightSample &s, const D3 &Ps) {
  const float sample_intensity = lt.intensity / (float)max(lt.sample_count, 1);      // Light::SetIntensity, fj_light.cc:38-50
  switch (lt.kind) {
    case 0: return c3(fmul(lt.intensity, lt.color[0]), fmul(lt.intensity, lt.color[1]), fmul(lt.intensity, lt.color[2]));
    case 1: {
      const D3 Ln = normalize(Ps - s.P);
      double dt = dot(Ln, s.N);
      if (lt.double_sided) dt = fabs(dt); else dt = dt > 0. ? dt : 0.;
      const float f = (float)dmul(dt, (double)sample_intensity);
      return c3(fmul(lt.color[0], f), fmul(lt.color[1], f), fmul(lt.color[2], f));
    }
    case 2: {
      const D3 Ln = normalize(Ps - s.P);
      if (dot(Ln, s.N) > 0) return c3(fmul(sample_intensity, lt.color[0]), fmul(sample_intensity, lt.color[1]), fmul(sample_intensity, lt.color[2]));
      return c3(0, 0, 0);
    }
    default: return c3(fmul(sample_intensity, s.color.r), fmul(sample_intensity, s.color.g), fmul(sample_intensity, s.color.b));
  }
}

// ------------------------------------------------------------------------------------------ shading
#define FJ_LIGHT_CACHE 16     // light samples a plastic hit keeps between deciding and writing its shadow-ray block
struct PathKey { uint32_t seed, tile, sample; };
struct ShadeCounters { unsigned int rays[5]; unsigned int hits, levels; };

// Shades one traced ray: adds its radiance (times the ray's throughput) to the sample through `sink.add`, and hands the
// secondary rays it spawns to `sink.spawn` (SlTrace -> trace_surface -> Shader::evaluate, src/fj_shading.cc:140-179,
// 527-572).  Sink: add(r,g,b) in float, alpha(a) for camera rays, spawn(const RayRec&).
//
// Shadow rays (SlIlluminance, src/fj_shading.cc:296-359): `Cl *= 1 - alpha_occluder` has to precede the Lambert weight and
// the float sum over the light samples runs in sample order, so the diffuse term of a plastic hit cannot be split into
// independent rays.  DEFER = false (megakernel): they are traced inline, on the binary BVH.  DEFER = true (wavefront): the
// hit writes ONE CONTIGUOUS BLOCK into the next queue — a header record (RAY_SHADOW_HEAD: parent throughput, shader slot,
// diffuse-map texel, sample count; tmin > tmax so k_extend retires it at the root) followed by one RAY_SHADOW record per
// light sample that passed the cone and intensity tests (ray, unshadowed Cl in `thr`, Lambert weight in `pad2`) — the
// block is traced by k_extend with everything else, and the next k_shade finishes the sum from the block in sample
// order (finish_shadow_block): the same float operations in the same order as the inline path, bit for bit.
template <typename T, bool PLASTIC = true, bool DEFER = false>
struct Shading {
  const DScene &sc; const DFrame &fr; PathKey key; ShadeCounters &cnt;
  __device__ Shading(const DScene &s, const DFrame &f, PathKey k, ShadeCounters &c) : sc(s), fr(f), key(k), cnt(c) {}
  __device__ __forceinline__ double rnd(unsigned long long node, uint32_t dim) const { return ctr_rand(key.seed, key.tile, key.sample, node, dim); }

  // SlIlluminance up to the shadow ray: direction and distance to the light sample, cone test, Light::Illuminate, the
  // 1e-4 intensity cut (src/fj_shading.cc:296-336).
  __device__ __forceinline__ bool light_sample_reaches(const DLight &lt, const LightSample &ls, const D3 &Ps, const D3 &axis,
                                                       C3 *Cl, D3 *Ln_out, double *dist) {
    D3 Ln = ls.P - Ps;
    const double distance = length(Ln);
    if (distance > 0) { const double inv = ddiv(1., distance); Ln = Ln * inv; }
    *Ln_out = Ln; *dist = distance;
    const D3 nml_axis = normalize(axis);
    const double cosangle = dot(nml_axis, Ln);
    if (cosangle < 6.123233995736766e-17) return false;          // cos(PI/2) in FP64
    const C3 lc = light_illuminate(lt, ls, Ps);
    if ((double)lc.r < .0001 && (double)lc.g < .0001 && (double)lc.b < .0001) return false;
    *Cl = lc;
    return true;
  }

  // The light samples of one shading point in the reference's order (SlNewLightSamples, fj_shading.cc:380-404;
  // Light::GetSamples of the four light types): f(light, sample) for every one of them.
  template <typename F>
  __device__ __forceinline__ void for_each_light_sample(unsigned long long node, F f) {
    uint32_t dim = 16;
    for (int li = 0; li < sc.nlights; li++) {
      const DLight &lt = sc.lights[li];
      const int ns = lt.kind == 0 ? 1 : (lt.kind == 3 ? min(lt.sample_count, lt.dome_count) : lt.sample_count);
      D3 gridN = mk(0, 0, 0);
      if (lt.kind == 1) gridN = normalize(mat_vector(lt.fwd, mk(0., 1., 0.)));
      for (int i = 0; i < ns; i++) {
        LightSample ls; ls.N = mk(0, 0, 0); ls.color = c3(0, 0, 0);
        if (lt.kind == 0) {                                        // fj_point_light.cc:21-39
          ls.P = mk(lt.translate[0], lt.translate[1], lt.translate[2]);
        } else if (lt.kind == 1) {                                 // fj_rectangle_light.cc:26-47
          double rx, rz;
          if (dim & 1u) { rx = rnd(node, dim); rz = rnd(node, dim + 1); }      // (odd only after a sphere light's 3-draw tries)
          else ctr_rand2(key.seed, key.tile, key.sample, node, dim, &rx, &rz);
          dim += 2;
          const double x = dsub(rx, .5), z = dsub(rz, .5);
          ls.P = mat_point(lt.fwd, mk(x, 0., z)); ls.N = gridN;
        } else if (lt.kind == 2) {                                 // fj_sphere_light.cc:22-46, XorShift::HollowSphereRand
          D3 p; double dd;
          for (;;) {
            p.x = dsub(dmul(2., rnd(node, dim)), 1.); p.y = dsub(dmul(2., rnd(node, dim + 1)), 1.); p.z = dsub(dmul(2., rnd(node, dim + 2)), 1.); dim += 3;
            dd = dot(p, p);
            if (dd > 0 && dd <= 1) break;
          }
          p = p * ddiv(1., __dsqrt_rn(dd));
          ls.P = mat_point(lt.fwd, p); ls.N = normalize(mat_vector(lt.fwd, p));
        } else {                                                   // fj_dome_light.cc:25-52
          const D3 dir = mk(lt.dome_dirs[3 * i], lt.dome_dirs[3 * i + 1], lt.dome_dirs[3 * i + 2]);
          ls.P = mat_point(lt.fwd, dir * (double)FLT_MAX);
          ls.N = mat_vector(lt.fwd, -1. * dir);
          ls.color = c3(lt.dome_colors[3 * i], lt.dome_colors[3 * i + 1], lt.dome_colors[3 * i + 2]);
        }
        f(lt, ls);
      }
    }
  }

  // diff = sum over all light samples of max(0, Nf.Ln) * Cl — the loop of PlasticShader::evaluate (plastic_shader.cc:117-137),
  // shadow rays traced inline (SlShadowContext :266-279).
  template <bool TRACE>
  __device__ C3 gather_lights(const D3 &P, const D3 &Nf, int shaded_object, unsigned long long node) {
    C3 diff = c3(0, 0, 0);
    for_each_light_sample(node, [&](const DLight &lt, const LightSample &ls) {
      C3 lc; D3 Ln; double distance;
      if (!light_sample_reaches(lt, ls, P, Nf, &lc, &Ln, &distance)) return;
      if (TRACE && fr.cast_shadow) {
        RayD sr; sr.o = P; sr.d = Ln; sr.tmin = .0001; sr.tmax = distance; sr.tidx = key.sample;
        Hit h;
        cnt.rays[RAY_SHADOW]++;
        if (trace_closest<T>(sc, sc.inst[shaded_object].shadow_target, sr, &h)) {
          cnt.hits++; cnt.levels += sc.meshes[sc.inst[h.inst].mesh].log2_tris;
          const float ac = fadd(1.f, -occluder_opacity(sc, h));
          lc.r = fmul(lc.r, ac); lc.g = fmul(lc.g, ac); lc.b = fmul(lc.b, ac);
        }
      }
      float Kd = (float)dot(Nf, Ln);
      Kd = Kd > 0.f ? Kd : 0.f;
      diff.r = fadd(diff.r, fmul(Kd, lc.r)); diff.g = fadd(diff.g, fmul(Kd, lc.g)); diff.b = fadd(diff.b, fmul(Kd, lc.b));
    });
    return diff;
  }

  // The wavefront's version: writes the block described above through `sink` and returns; the sum is finished one round later.
  template <typename Sink>
  __device__ void defer_lights(const RayRec &cur, const C3 &thr, const D3 &P, const D3 &Nf, int shaded_object, int slot, const float4 &dm, Sink &sink) {
    // one pass over the light samples decides which reach the point and keeps their rays (up to FJ_LIGHT_CACHE of them; a
    // second pass recomputes only what did not fit), then the block is reserved and written
    struct Kept { double lx, ly, lz, dist; float r, g, b, kd; };
    Kept kept[FJ_LIGHT_CACHE];
    int n = 0;
    for_each_light_sample(cur.node, [&](const DLight &lt, const LightSample &ls) {
      C3 lc; D3 Ln; double distance;
      if (!light_sample_reaches(lt, ls, P, Nf, &lc, &Ln, &distance)) return;
      if (n < FJ_LIGHT_CACHE) {
        float Kd = (float)dot(Nf, Ln);
        Kd = Kd > 0.f ? Kd : 0.f;
        Kept &k = kept[n]; k.lx = Ln.x; k.ly = Ln.y; k.lz = Ln.z; k.dist = distance; k.r = lc.r; k.g = lc.g; k.b = lc.b; k.kd = Kd;
      }
      n++;
    });
    if (n == 0) return;                                            // diff = 0: the diffuse term adds nothing
    unsigned at = sink.reserve((unsigned)n + 1u);
    RayRec c; memset(&c, 0, sizeof c);
    c.d[2] = 1.; c.tmin = 1.; c.tmax = 0.;                         // never enters a box: retired at the root of the instance tree
    c.thr[0] = thr.r; c.thr[1] = thr.g; c.thr[2] = thr.b; c.slot = cur.slot;
    c.node = (unsigned long long)(unsigned)n | ((unsigned long long)__float_as_uint(dm.x) << 32);
    c.pad2 = __float_as_int(dm.y); c.pad3 = __float_as_int(dm.z);
    c.target = sc.inst[shaded_object].shadow_target; c.type = RAY_SHADOW_HEAD; c.filter_shader = slot;
    sink.put(at++, c);
    RayRec r; memset(&r, 0, sizeof r);
    r.key = key.sample;                                            // shadow rays see the scene at the sample's time (SlShadowContext copies cxt)
    r.o[0] = P.x; r.o[1] = P.y; r.o[2] = P.z; r.tmin = .0001; r.slot = cur.slot;
    r.target = sc.inst[shaded_object].shadow_target; r.type = RAY_SHADOW; r.filter_shader = -1;
    const int nc = min(n, FJ_LIGHT_CACHE);
    for (int k = 0; k < nc; k++) {
      r.d[0] = kept[k].lx; r.d[1] = kept[k].ly; r.d[2] = kept[k].lz; r.tmax = kept[k].dist;
      r.thr[0] = kept[k].r; r.thr[1] = kept[k].g; r.thr[2] = kept[k].b; r.pad2 = __float_as_int(kept[k].kd);
      sink.put(at++, r);
    }
    if (n > FJ_LIGHT_CACHE) {
      int seen = 0;
      for_each_light_sample(cur.node, [&](const DLight &lt, const LightSample &ls) {
        C3 lc; D3 Ln; double distance;
        if (!light_sample_reaches(lt, ls, P, Nf, &lc, &Ln, &distance)) return;
        if (seen++ < FJ_LIGHT_CACHE) return;
        float Kd = (float)dot(Nf, Ln);
        Kd = Kd > 0.f ? Kd : 0.f;
        r.d[0] = Ln.x; r.d[1] = Ln.y; r.d[2] = Ln.z; r.tmax = distance;
        r.thr[0] = lc.r; r.thr[1] = lc.g; r.thr[2] = lc.b; r.pad2 = __float_as_int(Kd);
        sink.put(at++, r);
      });
    }
  }

  template <typename Sink>
  __device__ void shade(const RayRec &cur, const Hit &h, Sink &sink) {
    RayD ray; ray.o = mk(cur.o[0], cur.o[1], cur.o[2]); ray.d = mk(cur.d[0], cur.d[1], cur.d[2]); ray.tmin = cur.tmin; ray.tmax = cur.tmax;
    ray.tidx = key.sample;                                         // key.sample = y * nsx + x = the sample's entry of the time table
    cnt.hits++; cnt.levels += sc.meshes[sc.inst[h.inst].mesh].log2_tris;
    C3 thr = c3(cur.thr[0], cur.thr[1], cur.thr[2]);
    if (cur.filter_shader >= 0) {    // pathtracing_shader.cc:247-251: C *= pow(transmit, t_hit) of the refracted child
      const DShader &fs = sc.shaders[cur.filter_shader];
      thr.r = fmul(thr.r, (float)pow((double)fs.transmit[0], h.t));
      thr.g = fmul(thr.g, (float)pow((double)fs.transmit[1], h.t));
      thr.b = fmul(thr.b, (float)pow((double)fs.transmit[2], h.t));
    }
    D3 P, N; int slot;
    hit_surface(sc, ray, h, &P, &N, &slot);
    const DInstance &in = sc.inst[h.inst];
    float Os = 1.f;
    const int kind = slot < 0 ? 0 : sc.shaders[slot].kind;
    RayRec c;
    c.o[0] = P.x; c.o[1] = P.y; c.o[2] = P.z; c.tmax = 1000.; c.slot = cur.slot; c.key = key.sample; c.pad2 = c.pad3 = 0; c.filter_shader = -1;
    c.dd = cur.dd; c.rd = cur.rd; c.fd = cur.fd;
    if (kind == 0) {                                               // NO_SHADER_COLOR, fj_shading.cc:26,555-560
      sink.add(fmul(thr.r, .5f), thr.g, 0.f);
    } else if (kind == 1) {                                        // ConstantShader::evaluate, constant_shader.cc:72-94
      const DShader &sh = sc.shaders[slot];
      if (sh.texture) {                                            // C_tex = texture->Lookup(uv) * diffuse
        float tu, tv; hit_uv(sc, h, &tu, &tv);
        const float4 ct = tex_lookup(sc.textures[sh.texture - 1], tu, tv);
        sink.add(fmul(thr.r, fmul(ct.x, sh.diffuse[0])), fmul(thr.g, fmul(ct.y, sh.diffuse[1])), fmul(thr.b, fmul(ct.z, sh.diffuse[2])));
      } else
      sink.add(fmul(thr.r, sh.diffuse[0]), fmul(thr.g, sh.diffuse[1]), fmul(thr.b, sh.diffuse[2]));
    } else if (PLASTIC && kind == 2) {                             // PlasticShader::evaluate, plastic_shader.cc:101-179
      const DShader &sh = sc.shaders[slot];
      D3 Nf = sl_faceforward(ray.d, N);
      if (sh.bump_texture) {                                       // plastic_shader.cc:115-123
        float tu, tv; hit_uv(sc, h, &tu, &tv);
        D3 dPdu, dPdv; hit_derivatives(sc, h, key.sample, &dPdu, &dPdv);
        Nf = sl_bump_mapping(sc.textures[sh.bump_texture - 1], dPdu, dPdv, tu, tv, (double)sh.bump_amplitude, Nf);
      }
      float4 dm = make_float4(1.f, 1.f, 1.f, 1.f);                 // diffuse_map, plastic_shader.cc:148-156: Cs = diff * diffuse * diff_map
      if (sh.texture) { float tu, tv; hit_uv(sc, h, &tu, &tv); dm = tex_lookup(sc.textures[sh.texture - 1], tu, tv); }
      if (DEFER && fr.cast_shadow) defer_lights(cur, thr, P, Nf, h.inst, slot, dm, sink);
      else {
        const C3 diff = gather_lights<!DEFER>(P, Nf, h.inst, cur.node);      // (DEFER: only reached without shadow casting)
        sink.add(fmul(thr.r, fmul(fmul(diff.r, sh.diffuse[0]), dm.x)), fmul(thr.g, fmul(fmul(diff.g, sh.diffuse[1]), dm.y)), fmul(thr.b, fmul(fmul(diff.b, sh.diffuse[2]), dm.z)));
      }
      if (sh.do_reflect && (int)cur.rd + 1 <= fr.max_reflect) {    // SlReflectContext :242-252, gate :467-499
        const double Kr = sl_fresnel(ray.d, Nf, ddiv(1., (double)sh.ior));
        const D3 R = normalize(sl_reflect(ray.d, Nf));
        c.d[0] = R.x; c.d[1] = R.y; c.d[2] = R.z; c.tmin = .001;
        c.thr[0] = fmul(thr.r, (float)dmul(Kr, (double)sh.reflect[0])); c.thr[1] = fmul(thr.g, (float)dmul(Kr, (double)sh.reflect[1]));
        c.thr[2] = fmul(thr.b, (float)dmul(Kr, (double)sh.reflect[2]));
        c.node = cur.node * 4 + 2; c.target = in.reflect_target; c.type = RAY_REFLECT; c.rd = cur.rd + 1;
        sink.spawn(c);
        c.rd = cur.rd;
      }
      Os = sh.opacity;
    } else if (kind == 4) {                                        // GlassShader::evaluate, glass_shader.cc:88-133
      const DShader &sh = sc.shaders[slot];
      const double ior = ddiv(1., (double)sh.ior);
      const double Kr = sl_fresnel(ray.d, N, ior), Kt = dsub(1., Kr);
      if ((int)cur.rd + 1 <= fr.max_reflect) {                     // SlReflectContext + SlTrace(.0001, 1000); Cs += Kr * C_refl
        const D3 R = normalize(sl_reflect(ray.d, N));
        c.d[0] = R.x; c.d[1] = R.y; c.d[2] = R.z; c.tmin = .0001;
        c.thr[0] = fmul(thr.r, (float)Kr); c.thr[1] = fmul(thr.g, (float)Kr); c.thr[2] = fmul(thr.b, (float)Kr);
        c.node = cur.node * 4 + 2; c.target = in.reflect_target; c.type = RAY_REFLECT; c.rd = cur.rd + 1;
        sink.spawn(c);
        c.rd = cur.rd;
      }
      if ((int)cur.fd + 1 <= fr.max_refract) {                     // SlRefractContext; C_refr *= pow(filter_color, t_hit) when entering
        const D3 Tr = normalize(sl_refract(ray.d, N, ior));
        c.d[0] = Tr.x; c.d[1] = Tr.y; c.d[2] = Tr.z; c.tmin = .0001;
        c.thr[0] = fmul(thr.r, (float)Kt); c.thr[1] = fmul(thr.g, (float)Kt); c.thr[2] = fmul(thr.b, (float)Kt);
        c.filter_shader = (sh.do_color_filter && dot(ray.d, N) < 0) ? slot : -1;
        c.node = cur.node * 4 + 3; c.target = in.refract_target; c.type = RAY_REFRACT; c.fd = cur.fd + 1;
        sink.spawn(c);
      }
    } else {                                                       // PathtracingShader::evaluate, pathtracing_shader.cc:125-257
      const DShader &sh = sc.shaders[slot];
      if (sh.bump_texture) {                                       // :136-144: the integrators see the bumped normal (in_modified.N)
        float tu, tv; hit_uv(sc, h, &tu, &tv);
        D3 dPdu, dPdv; hit_derivatives(sc, h, key.sample, &dPdu, &dPdv);
        N = sl_bump_mapping(sc.textures[sh.bump_texture - 1], dPdu, dPdv, tu, tv, (double)sh.bump_amplitude, N);
      }
      sink.add(fmul(thr.r, sh.emission[0]), fmul(thr.g, sh.emission[1]), fmul(thr.b, sh.emission[2]));
      if (luminance(sh.diffuse) > 0.f && (int)cur.dd + 1 <= fr.max_diffuse) {       // integrate_diffuse :176-208
        const D3 w = N;
        D3 u = fabs(w.x) > .001 ? mk(0., 1., 0.) : mk(1., 0., 0.);
        u = normalize(cross(u, w));
        const D3 v = cross(w, u);
        const double x1 = rnd(cur.node, 0), x2 = rnd(cur.node, 1);
        const double r1 = dmul(dmul(2., 3.14159265358979323846), x1), r2 = x2, r2s = __dsqrt_rn(r2);
        const D3 D = normalize(((u * cos(r1)) * r2s + (v * sin(r1)) * r2s) + w * __dsqrt_rn(dsub(1., r2)));
        const float Kd = (float)dot(N, D);
        c.d[0] = D.x; c.d[1] = D.y; c.d[2] = D.z; c.tmin = .001;
        if (sh.texture) {                            // diffuse_map scales Cd (pathtracing_shader.cc:132-135): L = Cd * Kd * diffuse * C
          float tu, tv; hit_uv(sc, h, &tu, &tv);
          const float4 cd = tex_lookup(sc.textures[sh.texture - 1], tu, tv);
          c.thr[0] = fmul(thr.r, fmul(fmul(cd.x, Kd), sh.diffuse[0])); c.thr[1] = fmul(thr.g, fmul(fmul(cd.y, Kd), sh.diffuse[1])); c.thr[2] = fmul(thr.b, fmul(fmul(cd.z, Kd), sh.diffuse[2]));
        } else {
        c.thr[0] = fmul(thr.r, fmul(Kd, sh.diffuse[0])); c.thr[1] = fmul(thr.g, fmul(Kd, sh.diffuse[1])); c.thr[2] = fmul(thr.b, fmul(Kd, sh.diffuse[2]));
        }
        c.node = cur.node * 4 + 1; c.target = in.reflect_target; c.type = RAY_DIFFUSE; c.dd = cur.dd + 1;
        sink.spawn(c);
        c.dd = cur.dd;
      }
      if (luminance(sh.reflect) > 0.f && (int)cur.rd + 1 <= fr.max_reflect) {       // integrate_reflect :210-229
        const float kr = (float)sl_fresnel(ray.d, N, ddiv(1., (double)sh.ior));
        const D3 R = normalize(sl_reflect(ray.d, N));
        c.d[0] = R.x; c.d[1] = R.y; c.d[2] = R.z; c.tmin = .001;
        c.thr[0] = fmul(thr.r, fmul(kr, sh.reflect[0])); c.thr[1] = fmul(thr.g, fmul(kr, sh.reflect[1])); c.thr[2] = fmul(thr.b, fmul(kr, sh.reflect[2]));
        c.node = cur.node * 4 + 2; c.target = in.reflect_target; c.type = RAY_REFLECT; c.rd = cur.rd + 1;
        sink.spawn(c);
        c.rd = cur.rd;
      }
      if (luminance(sh.refract) > 0.f && (int)cur.fd + 1 <= fr.max_refract) {       // integrate_refract :231-257
        const double ior = ddiv(1., (double)sh.ior);
        const float kt = (float)dsub(1., sl_fresnel(ray.d, N, ior));
        const D3 Tr = normalize(sl_refract(ray.d, N, ior));
        c.d[0] = Tr.x; c.d[1] = Tr.y; c.d[2] = Tr.z; c.tmin = .0001;
        c.thr[0] = fmul(thr.r, fmul(kt, sh.refract[0])); c.thr[1] = fmul(thr.g, fmul(kt, sh.refract[1])); c.thr[2] = fmul(thr.b, fmul(kt, sh.refract[2]));
        c.filter_shader = (sh.do_color_filter && dot(ray.d, N) < 0) ? slot : -1;
        c.node = cur.node * 4 + 3; c.target = in.refract_target; c.type = RAY_REFRACT; c.fd = cur.fd + 1;
        sink.spawn(c);
      }
    }
    if (cur.type == RAY_CAMERA) sink.alpha(fminf(fmaxf(Os, 0.f), 1.f));               // trace_surface :562-566
  }
};

// Sample (x, y) and tile of a slot — inverse of sample_slot().
__device__ __forceinline__ void slot_decode(const DFrame &fr, const DTile *tiles, uint32_t wstride, uint32_t slot, int *ti, TileGrid *g, int *x, int *y) {
  *ti = (int)(slot / wstride);
  const uint32_t r = slot % wstride, blk = r >> 5, l = r & 31;
  *g = tile_grid(fr, tiles[*ti]);
  *x = (int)(blk % g->nbx) * 8 + (int)(l & 7); *y = (int)(blk / g->nbx) * 4 + (int)(l >> 3);
}

// ------------------------------------------------------------------------------------------ frame arguments
struct QueueCtl { unsigned int count[2]; unsigned int head; unsigned int overflow; };
struct RenderArgs {
  DScene sc; DCamera cam; DFrame fr;
  const DTile *tiles; int ntiles;         // tiles of this batch
  uint32_t wstride;                       // sample slots per tile (multiple of 32)
  Accum *accum;                           // ntiles * wstride per-sample accumulators
  DCounters *counters;
  unsigned long long *work;               // megakernel: global work counter (units of 32 slots)
  RayRec *queue[2]; HitRec *hits; QueueCtl *ctl; uint32_t capacity; int cur;     // wavefront
  int refill, phase_a_min, park;          // k_extend scheduling: refill / phase-A thresholds (lanes), speculative leaf parking
  const char *top_src; int top_count;     // k_extend2<TOP>: the NodeQ64 array whose first top_count nodes are staged in shared memory
  int shadow_anyhit;                      // k_extend2: every shader of the scene is opaque -> shadow rays stop at their first hit
  int shade_prefetch;                     // k_shade: prefetch the warp's next 32 records into L2 (FJGPU_SHADE_PREFETCH)
  int b1_min, b2_min;                     // k_extend_ring: (ray, triangle) pairs / entering lanes that make a heavy phase worth running
  int chunked;                            // k_shade without plastic shaders: warps reserve queue slots in chunks (QueueSink)
  // ray sorting between bounces: counting sort of the next queue by (direction octant | origin cell)
  unsigned int *hist;                     // sort_bins + 1 counters (null = no sorting)
  unsigned int *perm;                     // order in which k_extend walks queue[cur] (null = queue order)
  float sort_lo[3], sort_scale[3]; int sort_bits; unsigned int sort_bins;
};

__device__ __forceinline__ void flush_counters(const ShadeCounters &c, unsigned long long nsamp, DCounters *out, int lane) {
  unsigned long long v[8] = {c.rays[0], c.rays[1], c.rays[2], c.rays[3], c.rays[4], c.hits, c.levels, nsamp};
  for (int k = 0; k < 8; k++) {
    unsigned long long x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0 && x) atomicAdd(k < 5 ? &out->rays[k] : (k == 5 ? &out->hits : (k == 6 ? &out->levels : &out->samples)), x);
  }
}

// ------------------------------------------------------------------------------------------ megakernel (cross-check path)
// One camera sample per lane with a private DFS stack of pending rays; warps pull 8x4 sample blocks from a global
// counter.  Kept as the independent second implementation the parity tests compare the wavefront against
// (FJGPU_FLAG_MEGAKERNEL), with FP32 or FP64 (FJGPU_FLAG_FP64_BOXES) box culling.
struct StackSink {
  RayRec *stack; int sp; C3 acc; float a;
  __device__ __forceinline__ void add(float r, float g, float b) { acc.r = fadd(acc.r, r); acc.g = fadd(acc.g, g); acc.b = fadd(acc.b, b); }
  __device__ __forceinline__ void alpha(float v) { a = v; }
  __device__ __forceinline__ void spawn(const RayRec &c) { if (sp < FJ_PENDING) stack[sp++] = c; }
  __device__ __forceinline__ unsigned reserve(unsigned) { return 0u; }      // (the megakernel traces shadow rays inline)
  __device__ __forceinline__ void put(unsigned, const RayRec &) {}
};

template <typename T>
__global__ void __launch_bounds__(128) k_render_samples(const RenderArgs a) {
  const int lane = threadIdx.x & 31;
  const unsigned long long total = (unsigned long long)a.ntiles * a.wstride;
  ShadeCounters cnt; memset(&cnt, 0, sizeof cnt);
  unsigned long long nsamp = 0;
  RayRec stack[FJ_PENDING];
  for (;;) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.work, 32ull);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= total) break;
    const int ti = (int)(base / a.wstride);
    const uint32_t blk = (uint32_t)(base % a.wstride) >> 5;
    const DTile tile = a.tiles[ti];
    const TileGrid g = tile_grid(a.fr, tile);
    if (blk >= (uint32_t)(g.nbx * g.nby)) continue;
    const int x = (int)(blk % g.nbx) * 8 + (lane & 7), y = (int)(blk / g.nbx) * 4 + (lane >> 3);
    Accum out; out.r = out.g = out.b = 0; out.a = 0.f; out.pad = 0;
    if (x < g.nsx && y < g.nsy) {
      double u, v; sample_uv(a.fr, g, x, y, &u, &v);
      RayD ray; camera_ray(a.cam, u, v, (uint32_t)(y * g.nsx + x), &ray);
      PathKey key; key.seed = a.fr.seed; key.tile = (uint32_t)tile.id; key.sample = (uint32_t)(y * g.nsx + x);
      Shading<T> sh(a.sc, a.fr, key, cnt);
      StackSink sink; sink.stack = stack; sink.sp = 0; sink.acc = c3(0, 0, 0); sink.a = 0.f;
      RayRec cur;
      cur.o[0] = ray.o.x; cur.o[1] = ray.o.y; cur.o[2] = ray.o.z; cur.d[0] = ray.d.x; cur.d[1] = ray.d.y; cur.d[2] = ray.d.z;
      cur.tmin = ray.tmin; cur.tmax = ray.tmax; cur.thr[0] = cur.thr[1] = cur.thr[2] = 1.f; cur.slot = 0; cur.node = 1;
      cur.target = a.fr.target_group; cur.type = RAY_CAMERA; cur.dd = cur.rd = cur.fd = 0; cur.filter_shader = -1; cur.key = key.sample; cur.pad2 = cur.pad3 = 0;
      sink.spawn(cur);
      while (sink.sp > 0) {
        cur = stack[--sink.sp];
        RayD r; r.o = mk(cur.o[0], cur.o[1], cur.o[2]); r.d = mk(cur.d[0], cur.d[1], cur.d[2]); r.tmin = cur.tmin; r.tmax = cur.tmax; r.tidx = key.sample;
        cnt.rays[cur.type]++;
        Hit h;
        if (!trace_closest<T>(a.sc, cur.target, r, &h)) continue;
        sh.shade(cur, h, sink);
      }
      out.r = to_fix(sink.acc.r); out.g = to_fix(sink.acc.g); out.b = to_fix(sink.acc.b); out.a = sink.a;
      nsamp++;
    }
    a.accum[(size_t)ti * a.wstride + (blk << 5) + lane] = out;
  }
  flush_counters(cnt, nsamp, a.counters, lane);
}

// ------------------------------------------------------------------------------------------ wavefront: generate
// Sampler + camera: one camera ray per sample of the batch into queue[cur]; zeroes the accumulators
// (FixedGridSampler::generate_samples + Camera::GetRay, integrate_samples loop head src/fj_renderer.cc:1061-1075).
__global__ void __launch_bounds__(256) k_generate(const RenderArgs a) {
  const int lane = threadIdx.x & 31;
  const unsigned long long total = (unsigned long long)a.ntiles * a.wstride;
  RayRec *q = a.queue[a.cur];
  unsigned long long nsamp = 0;
  for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w - lane < total; w += (unsigned long long)gridDim.x * blockDim.x) {
    bool valid = false; RayRec r;
    if (w < total) {
      int ti, x, y; TileGrid g;
      slot_decode(a.fr, a.tiles, a.wstride, (uint32_t)w, &ti, &g, &x, &y);
      Accum z; z.r = z.g = z.b = 0; z.a = 0.f; z.pad = 0;
      a.accum[w] = z;
      if (x < g.nsx && y < g.nsy) {
        double u, v; sample_uv(a.fr, g, x, y, &u, &v);
        RayD ray; camera_ray(a.cam, u, v, (uint32_t)(y * g.nsx + x), &ray);
        r.o[0] = ray.o.x; r.o[1] = ray.o.y; r.o[2] = ray.o.z; r.d[0] = ray.d.x; r.d[1] = ray.d.y; r.d[2] = ray.d.z;
        r.tmin = ray.tmin; r.tmax = ray.tmax; r.thr[0] = r.thr[1] = r.thr[2] = 1.f; r.slot = (uint32_t)w; r.node = 1;
        r.target = a.fr.target_group; r.type = RAY_CAMERA; r.dd = r.rd = r.fd = 0; r.filter_shader = -1; r.key = ray.tidx; r.pad2 = r.pad3 = 0;
        valid = true; nsamp++;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (m) {
      unsigned base = 0;
      const int leader = __ffs(m) - 1;
      if (lane == leader) base = atomicAdd(&a.ctl->count[a.cur], __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (valid) {
        const unsigned i = base + __popc(m & ((1u << lane) - 1));
        if (i < a.capacity) store_ray_cs(q + i, r); else a.ctl->overflow = 1;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) nsamp += __shfl_down_sync(0xffffffffu, nsamp, o);
  if (lane == 0 && nsamp) atomicAdd(&a.counters->samples, nsamp);
}

// ------------------------------------------------------------------------------------------ wavefront: extend
// Persistent-threads closest-hit kernel over queue[cur]: every lane owns one ray at a time and pulls the next one from
// the queue head as soon as a quarter of the warp has run dry (warp-aggregated atomic), so the traversal loop keeps
// >= 24 lanes busy instead of waiting for the slowest ray of a fixed batch.  Semantics of trace_closest():
// Accelerator::Intersect on the group's accelerator (src/fj_shading.cc:538-541).
//
// Box culling is FP32 with a rigorous error bound folded into the FMA constants (no per-node slack arithmetic):
//   t(plane) = fma(plane, 1/d, -(o/d) -+ e),  e = 2^-20 (|o| + B) |1/d|  >=  every rounding of o, 1/d, the product and the
//   fma for planes with |plane| <= B (B = bound magnitude of the space being traversed), so a box the exact FP64 ray
//   touches within [tmin, best_t] is never culled.  Triangles are then tested in exact FP64 (tri_intersect).
struct BoxRay32 { float ix, iy, iz, nx, ny, nz, fx, fy, fz; };
__device__ __forceinline__ int box_octant(const BoxRay32 &r) { return (r.ix < 0.f ? 1 : 0) | (r.iy < 0.f ? 2 : 0) | (r.iz < 0.f ? 4 : 0); }
__device__ __forceinline__ void box_axis(double o, double d, float B, float *inv, float *cn, float *cf) {
  float df = (float)d;
  if (!(fabsf(df) >= 1e-18f)) df = (df < 0.f || (df == 0.f && signbit(df))) ? -1e-18f : 1e-18f;
  const float i = __frcp_rn(df);
  const float of = (float)o;
  const float p = __fmul_rn(of, i);
  const float e = __fmul_rn(__fmul_rn(__fadd_rn(fabsf(of), B), fabsf(i)), 9.5367431640625e-07f);
  *inv = i; *cn = __fsub_rn(-p, e); *cf = __fadd_rn(-p, e);
}
__device__ __forceinline__ void make_box_ray32(const D3 &o, const D3 &d, float B, BoxRay32 &r) {
  box_axis(o.x, d.x, B, &r.ix, &r.nx, &r.fx);
  box_axis(o.y, d.y, B, &r.iy, &r.ny, &r.fy);
  box_axis(o.z, d.z, B, &r.iz, &r.nz, &r.fz);
}

// Warp-synchronous structure (every loop is controlled by a ballot, so the 32 lanes stay converged by construction):
//   refill   when >= FJ_REFILL lanes are idle, they take the next rays from the queue head (one warp-aggregated atomic)
//   phase A  inner-node steps for all lanes that have one; a lane that reaches a triangle leaf parks it in `leaf`
//            and keeps descending with the next stack entry (speculative traversal) until it holds a second leaf
//   phase B  the parked leaves: one exact FP64 triangle test per lane per iteration; then the rare transitions
//            (enter an instance, leave it, finish the ray and write its hit record)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_extend(const RenderArgs a) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const RayRec *rays = a.queue[a.cur];
  if (blockIdx.x == 0 && threadIdx.x == 0) a.ctl->count[a.cur ^ 1] = 0;      // the queue k_shade fills next
  const unsigned count = min(a.ctl->count[a.cur], a.capacity);
  const DScene &sc = a.sc;
  int stack[FJ_STACK4];
  const int SENTINEL = (int)0x80000000, DONE = (int)0x80000001;
  const unsigned MISS = 0xffffffffu;

  bool active = false, drained = false;
  unsigned ridx = 0, n_steps = 0, n_tris = 0;
  int sp = 0, node = DONE, leaf = 0, cur_inst = -1, best_prim = -1, best_inst = -1;
  bool in_blas = false, found = false, br_world = false;
  double tmin = 0, tmax = 0, best_t = 0, best_u = 0, best_v = 0;
  float tn = 0, tf = 0;
  D3 o = mk(0, 0, 0), d = mk(0, 0, 0);          // object-space ray while inside a BLAS
  BoxRay32 br; br.ix = br.iy = br.iz = br.nx = br.ny = br.nz = br.fx = br.fy = br.fz = 0.f;
  int oct = 0;                                  // direction signs of the current box ray (bit a: 1/d[a] < 0)
  const float4 *nodes = nullptr, *tlas = nullptr; const int32_t *order = nullptr; float tlas_B = 0;
  const float4 *tri32 = nullptr; const double *tri64 = nullptr, *tri64v = nullptr;

  for (;;) {
    // ---- refill idle lanes from the queue head
    const unsigned idle = __ballot_sync(FULL, !active);
    if (idle == FULL && drained) break;
    if (!drained && __popc(idle) >= a.refill) {
      const int n = __popc(idle), leader = __ffs(idle) - 1;
      unsigned base = 0;
      if (lane == leader) base = atomicAdd(&a.ctl->head, (unsigned)n);
      base = __shfl_sync(FULL, base, leader);
      if (base + n >= count) drained = true;
      if (!active) {
        const unsigned i = base + __popc(idle & ((1u << lane) - 1));
        if (i < count) {
          ridx = a.perm ? a.perm[i] : i;
          const RayRec &r = rays[ridx];
          tmin = r.tmin; tmax = r.tmax; best_t = tmax; found = false; best_prim = -1; best_inst = -1; best_u = best_v = 0;
          tn = __double2float_rd(tmin); tf = __double2float_ru(best_t);
          const DGroup grp = sc.groups[r.target];
          tlas = grp.nodes4; order = grp.order; tlas_B = grp.bmag;
          nodes = tlas; in_blas = false; sp = 0; node = 0; leaf = 0;
          make_box_ray32(mk(r.o[0], r.o[1], r.o[2]), mk(r.d[0], r.d[1], r.d[2]), tlas_B, br); oct = box_octant(br);
          br_world = true;
          active = true;
        }
      }
      if (__ballot_sync(FULL, active) == 0) break;
    }

    // ---- phase A: inner nodes (both children, nearer first)
    for (;;) {
      const bool want = active && node >= 0 && (a.park || leaf == 0);
      const unsigned wm = __ballot_sync(FULL, want);
      if (wm == 0) break;
      // few lanes left descending: switch to the parked leaves / transitions if there are any to work on
      if (__popc(wm) < a.phase_a_min && __any_sync(FULL, active && (leaf != 0 || node < 0))) break;
      if (want) {
        if (!in_blas && !br_world) {             // back in the instance tree after a BLAS: rebuild the world-space box ray
          const RayRec &r = rays[ridx];
          make_box_ray32(mk(r.o[0], r.o[1], r.o[2]), mk(r.d[0], r.d[1], r.d[2]), tlas_B, br); oct = box_octant(br);
          br_world = true;
        }
        // 4-wide node: lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4] child[4].  The near / far plane of every axis is
        // picked by ADDRESS from the ray's direction signs (oct bit a set: hi is the near plane of axis a), not by selects.
        const float4 *np = nodes + 8 * (size_t)node;
        const int ox_ = oct & 1, oy_ = (oct >> 1) & 1, oz_ = (oct >> 2) & 1;
        const float4 nxp = __ldg(np + ox_), fxp = __ldg(np + (ox_ ^ 1)), nyp = __ldg(np + 2 + oy_), fyp = __ldg(np + 2 + (oy_ ^ 1));
        const float4 nzp = __ldg(np + 4 + oz_), fzp = __ldg(np + 4 + (oz_ ^ 1));
        const int4 ch = __ldg((const int4 *)(np + 6));
        unsigned key0, key1, key2, key3;
#define FJ_CHILD(KEY, K, C)                                                                                                     \
        {                                                                                                                          \
          const float nr = fmaxf(fmaxf(fmaf(nxp.C, br.ix, br.nx), fmaf(nyp.C, br.iy, br.ny)), fmaxf(fmaf(nzp.C, br.iz, br.nz), tn));  \
          const float fr_ = fminf(fminf(fmaf(fxp.C, br.ix, br.fx), fmaf(fyp.C, br.iy, br.fy)), fminf(fmaf(fzp.C, br.iz, br.fz), tf)); \
          KEY = nr <= fr_ ? ((__float_as_uint(nr) & ~3u) | K) : MISS;                                                              \
        }
        FJ_CHILD(key0, 0u, x) FJ_CHILD(key1, 1u, y) FJ_CHILD(key2, 2u, z) FJ_CHILD(key3, 3u, w)
#undef FJ_CHILD
        // entry distances are positive (tn > 0), so their bit patterns order like unsigned integers.  The nearest hit child
        // is visited next; the other hit children are pushed (exact front-to-back order for up to two hits, slot order beyond)
        const unsigned kmin = min(min(key0, key1), min(key2, key3));
        if (kmin != MISS) {
          const unsigned w = kmin & 3u;
          if (key0 != MISS && w != 0u) stack[sp++] = ch.x;
          if (key1 != MISS && w != 1u) stack[sp++] = ch.y;
          if (key2 != MISS && w != 2u) stack[sp++] = ch.z;
          if (key3 != MISS && w != 3u) stack[sp++] = ch.w;
          node = (w & 2u) ? ((w & 1u) ? ch.w : ch.z) : ((w & 1u) ? ch.y : ch.x);
        } else node = sp > 0 ? stack[--sp] : DONE;
        n_steps++;
        if (a.park && node < 0 && in_blas && node != SENTINEL && leaf == 0) { leaf = node; node = stack[--sp]; }     // SENTINEL is below every BLAS entry
      }
    }

    // ---- phase B1: parked triangle leaves, exact FP64 tests, one triangle per lane per iteration
    {
      const int ref = ~leaf;
      const int first = ref >> 3, cnt = leaf != 0 ? (ref & 7) + 1 : 0;
      for (int k = 0;; k++) {
        const bool want = k < cnt;
        if (!__any_sync(FULL, want)) break;
        if (want) {
          n_tris++;
          D3 v0, v1, v2; int prim;
          if (tri32) {
            const float4 *tp = tri32 + 3 * (size_t)(first + k);
            const float4 p0 = __ldg(tp), p1 = __ldg(tp + 1), p2 = __ldg(tp + 2);
            v0 = mk(p0.x, p0.y, p0.z); v1 = mk(p1.x, p1.y, p1.z); v2 = mk(p2.x, p2.y, p2.z); prim = __float_as_int(p0.w);
          } else if (tri64v) {                                       // `P0 += time * velocity0`, src/fj_mesh.cc:252-259
            const double *p = tri64v + 20 * (size_t)(first + k);
            const double tm = sc.time_tab[rays[ridx].key];
            v0 = mk(p[0], p[1], p[2]) + tm * mk(p[10], p[11], p[12]); v1 = mk(p[3], p[4], p[5]) + tm * mk(p[13], p[14], p[15]);
            v2 = mk(p[6], p[7], p[8]) + tm * mk(p[16], p[17], p[18]); prim = (int)__double_as_longlong(p[9]);
          } else {
            const double *p = tri64 + 10 * (size_t)(first + k);
            v0 = mk(p[0], p[1], p[2]); v1 = mk(p[3], p[4], p[5]); v2 = mk(p[6], p[7], p[8]); prim = (int)__double_as_longlong(p[9]);
          }
          double t, u, v;
          if (tri_intersect(v0, v1, v2, o, d, &t, &u, &v) && tmin <= t && t <= tmax) {       // RayInRange, src/fj_ray.h:29-32
            const bool better = !found ? true : (t < best_t || (t == best_t && (cur_inst < best_inst || (cur_inst == best_inst && prim > best_prim))));
            if (better) { found = true; best_t = t; best_u = u; best_v = v; best_prim = prim; best_inst = cur_inst; tf = __double2float_ru(t); }
          }
        }
      }
      leaf = 0;
    }

    // ---- phase B2: transitions of lanes whose next stack entry is not an inner node
    const bool special = active && node < 0;
    if (__any_sync(FULL, special)) {
      if (special) {
        if (node == DONE) {                        // traversal finished: write the hit record
          HitRec hr; hr.t = found ? best_t : FJ_REAL_MAX; hr.u = best_u; hr.v = best_v; hr.prim = best_prim; hr.inst = found ? best_inst : -1;
          store_hit_cs(a.hits + ridx, hr);
          active = false;
        } else if (node == SENTINEL) {             // the instance's BLAS is done: back to the instance tree (its box ray is
          in_blas = false; nodes = tlas; br_world = false;      // rebuilt only if an inner node of that tree is still to be visited)
          node = sp > 0 ? stack[--sp] : DONE;
        } else if (in_blas) {                      // a second triangle leaf: park it now that the slot is free
          leaf = node; node = stack[--sp];
        } else {                                   // TLAS leaf: enter the first instance, re-queue the others
          const int ref = ~node;
          const int first = ref >> 3, cnt = (ref & 7) + 1;
          for (int k = cnt - 1; k >= 1; k--) stack[sp++] = ~(((first + k) << 3) | 0);
          cur_inst = order[first];
          const DInstance &in = sc.inst[cur_inst];
          const RayRec &r = rays[ridx];
          const double *inv = inst_inv(in, r.key);
          o = mat_point(inv, mk(r.o[0], r.o[1], r.o[2]));
          d = mat_vector(inv, mk(r.d[0], r.d[1], r.d[2]));
          const DMesh &m = sc.meshes[in.mesh];
          make_box_ray32(o, d, m.bmag, br); oct = box_octant(br);
          br_world = false;
          nodes = m.nodes4; tri32 = m.tri32; tri64 = m.tri64; tri64v = m.tri64v;
          in_blas = true;
          stack[sp++] = SENTINEL;
          node = 0;
        }
      }
    }
  }
  // traversal statistics (4-wide node steps and exact triangle tests) for DESIGN.md / bench.py
  unsigned long long ns = n_steps, nt = n_tris;
  for (int off = 16; off > 0; off >>= 1) { ns += __shfl_down_sync(FULL, ns, off); nt += __shfl_down_sync(FULL, nt, off); }
  if (lane == 0 && a.counters) { atomicAdd(&a.counters->node_steps, ns); atomicAdd(&a.counters->tri_tests, nt); }
}

// ------------------------------------------------------------------------------------------ wavefront: shade
// Chunked slot reservation (RenderArgs::chunked): instead of one atomic on the queue counter per spawn, a warp takes
// FJ_QCHUNK slots at a time and hands them out through a counter in shared memory.  The warp's local positions [0, alloc)
// map to base[(p / FJ_QCHUNK) & 1] + p % FJ_QCHUNK; k_shade tops the reservation up at the head of every iteration so that
// FJ_QRESERVE slots (32 lanes x at most 4 children) are free, which keeps at most two chunks live.  What a warp has not
// used when it leaves the kernel is filled with RAY_DEAD records (tmin > tmax: k_extend retires them at the root, k_shade
// skips them) — the queue stays one dense range [0, count).
#define FJ_QCHUNK 256u
#define FJ_QRESERVE 128u
struct WarpChunks { unsigned used, alloc, base[2]; };

struct QueueSink {
  const RenderArgs &a; RayRec *next; int lane;
  long long r, g, b; bool has_alpha; float av;
  WarpChunks *wc;                         // null: one atomic on the queue counter per spawn
  __device__ __forceinline__ void add(float x, float y, float z) { r += to_fix(x); g += to_fix(y); b += to_fix(z); }
  __device__ __forceinline__ void alpha(float v) { has_alpha = true; av = v; }
  // writes record `c` into slot i of the next queue (and files it under its sort key when rays are sorted between bounces)
  __device__ __forceinline__ void put(unsigned i, const RayRec &c) {
    if (i < a.capacity) {
      store_ray_cs(next + i, c);
      if (a.hist) {       // sort key: rays leaving the same cell of the scene in the same octant walk the same part of the BVH
        const unsigned cells = (1u << a.sort_bits) - 1u;
        const unsigned cx = min((unsigned)fmaxf(((float)c.o[0] - a.sort_lo[0]) * a.sort_scale[0], 0.f), cells);
        const unsigned cy = min((unsigned)fmaxf(((float)c.o[1] - a.sort_lo[1]) * a.sort_scale[1], 0.f), cells);
        const unsigned cz = min((unsigned)fmaxf(((float)c.o[2] - a.sort_lo[2]) * a.sort_scale[2], 0.f), cells);
        unsigned mort = 0;
        for (int b = 0; b < a.sort_bits; b++) mort |= (((cx >> b) & 1u) << (3 * b)) | (((cy >> b) & 1u) << (3 * b + 1)) | (((cz >> b) & 1u) << (3 * b + 2));
        const unsigned oct = (c.d[0] < 0. ? 1u : 0u) | (c.d[1] < 0. ? 2u : 0u) | (c.d[2] < 0. ? 4u : 0u);
        const unsigned key = (oct << (3 * a.sort_bits)) | mort;
        next[i].key = key;
        atomicAdd(&a.hist[key], 1u);
      }
    } else a.ctl->overflow = 1;
  }
  // the warp may be diverged here: aggregate over whichever lanes arrive together
  __device__ __forceinline__ void spawn(const RayRec &c) {
    const unsigned m = __activemask();
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (wc) {
      if (lane == leader) base = atomicAdd(&wc->used, (unsigned)__popc(m));
      const unsigned p = __shfl_sync(m, base, leader) + __popc(m & ((1u << lane) - 1));
      put(wc->base[(p / FJ_QCHUNK) & 1u] + (p % FJ_QCHUNK), c);
      return;
    }
    if (lane == leader) base = atomicAdd(&a.ctl->count[a.cur ^ 1], (unsigned)__popc(m));
    base = __shfl_sync(m, base, leader);
    put(base + __popc(m & ((1u << lane) - 1)), c);
  }
  // `n` (<= 1023) contiguous slots for this lane; one atomic for the lanes that arrive together (prefix sums by bit plane)
  __device__ __forceinline__ unsigned reserve(unsigned n) {
    const unsigned m = __activemask(), lt = (1u << lane) - 1u;
    unsigned before = 0, total = 0;
#pragma unroll
    for (int b = 0; b < 10; b++) {
      const unsigned plane = __ballot_sync(m, (n >> b) & 1u);
      before += (unsigned)__popc(plane & lt) << b; total += (unsigned)__popc(plane) << b;
    }
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(&a.ctl->count[a.cur ^ 1], total);
    base = __shfl_sync(m, base, leader);
    return base + before;
  }
};

// Counting sort of queue[cur] by RayRec::key: exclusive scan of the histogram k_shade filled, then one pass that hands
// every ray a position inside its bin (order inside a bin is arbitrary and does not influence any result).
__global__ void __launch_bounds__(1024) k_sort_scan(unsigned int *hist, unsigned int bins) {
  __shared__ unsigned int part[1024];
  const unsigned per = (bins + 1023) / 1024, b0 = threadIdx.x * per;
  unsigned sum = 0;
  for (unsigned k = 0; k < per && b0 + k < bins; k++) sum += hist[b0 + k];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const unsigned v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned run = part[threadIdx.x] - sum;
  for (unsigned k = 0; k < per && b0 + k < bins; k++) { const unsigned c = hist[b0 + k]; hist[b0 + k] = run; run += c; }
}
__global__ void __launch_bounds__(256) k_sort_scatter(const RenderArgs a) {
  const unsigned count = min(a.ctl->count[a.cur], a.capacity);
  const RayRec *q = a.queue[a.cur];
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    a.perm[atomicAdd(&a.hist[q[i].key], 1u)] = i;
}

// One thread per traced ray of queue[cur]: shader evaluation, radiance into the sample accumulators, secondary rays
// compacted into queue[cur ^ 1] by warp ballot (Shader::Evaluate + the Sl*Context/SlTrace calls the plugins make).
// PLASTIC = false drops the light loop and its inline shadow-ray traversal from the kernel (scenes without a plastic
// shader): fewer registers, more resident warps.
// The term one traced shadow ray contributes to its block's sum: Kd * (Cl * (1 - alpha_occluder)), the float operations of
// Shading::gather_lights in the same order.
__device__ __forceinline__ C3 shadow_term(const DScene &sc, const float thr[3], int kdbits, const HitRec &hr) {
  C3 lc = c3(thr[0], thr[1], thr[2]);
  if (hr.inst >= 0) {
    Hit h; h.t = hr.t; h.u = hr.u; h.v = hr.v; h.prim = hr.prim; h.inst = hr.inst;
    const float ac = fadd(1.f, -occluder_opacity(sc, h));
    lc.r = fmul(lc.r, ac); lc.g = fmul(lc.g, ac); lc.b = fmul(lc.b, ac);
  }
  const float Kd = __int_as_float(kdbits);
  return c3(fmul(Kd, lc.r), fmul(Kd, lc.g), fmul(Kd, lc.b));
}

// One thread per traced ray of queue[cur]: shader evaluation, radiance into the sample accumulators, secondary rays
// compacted into queue[cur ^ 1] by warp ballot (Shader::Evaluate + the Sl*Context/SlTrace calls the plugins make).
// PLASTIC = false drops the light loop and the shadow-ray blocks from the kernel (scenes without a plastic shader).
// PLASTIC = true: a warp walks 32 consecutive records; the lanes that hold traced shadow rays compute their terms, the
// lane that holds the block's header gathers them IN SAMPLE ORDER with shuffles (and reads the few that lie beyond the
// warp's 32 records from memory), so the float sum equals the inline loop's bit for bit.
template <typename T, bool PLASTIC, int MINB = (PLASTIC ? 4 : 5)>
__global__ void __launch_bounds__(128, MINB) k_shade(const RenderArgs a) {
  const int lane = threadIdx.x & 31;
  const unsigned count = min(a.ctl->count[a.cur], a.capacity);
  if (blockIdx.x == 0 && threadIdx.x == 0) a.ctl->head = 0;                  // the next k_extend starts at the queue head
  const RayRec *rays = a.queue[a.cur];
  ShadeCounters cnt; memset(&cnt, 0, sizeof cnt);
  __shared__ WarpChunks chunks[4];
  WarpChunks *wc = (!PLASTIC && a.chunked) ? &chunks[threadIdx.x >> 5] : nullptr;
  if (wc && lane == 0) { wc->used = wc->alloc = 0; wc->base[0] = wc->base[1] = 0; }
  __syncwarp();
  // every warp walks whole 32-record groups (the loop bound is the same for its 32 lanes: the PLASTIC gather and the chunk
  // top-up below are warp-synchronous)
  const unsigned start = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u;
  for (unsigned i0 = start; i0 < count; i0 += gridDim.x * blockDim.x) {
    const unsigned i = i0 + lane;
    const bool valid = i < count;
    {                                       // the records of the warp's NEXT group on their way into L2 while this one is shaded: the
      const unsigned inext = i0 + gridDim.x * blockDim.x;      // kernel waits on these loads more than on anything else (ncu: long scoreboard)
      if (a.shade_prefetch && inext < count) {
        const unsigned nrec = min(32u, count - inext);
        const char *rb = reinterpret_cast<const char *>(rays + inext), *hb = reinterpret_cast<const char *>(a.hits + inext);
        if (128u * lane < nrec * (unsigned)sizeof(RayRec)) prefetch_l2(rb + 128u * lane);      // 32 x 112 B = 28 lines
        if (128u * lane < nrec * (unsigned)sizeof(HitRec)) prefetch_l2(hb + 128u * lane);      // 32 x 32 B = 8 lines
      }
    }
    if (wc) {                               // top the warp's slot reservation up to FJ_QRESERVE free slots
      __syncwarp();                         // every lane has finished the spawns of the previous iteration (they bump wc->used)
      if (lane == 0 && wc->alloc - wc->used < FJ_QRESERVE) {
        wc->base[(wc->alloc / FJ_QCHUNK) & 1u] = atomicAdd(&a.ctl->count[a.cur ^ 1], FJ_QCHUNK);
        wc->alloc += FJ_QCHUNK;
      }
      __syncwarp();
    }
    RayRec cur; HitRec hr; hr.inst = -1;
    if constexpr (PLASTIC) { cur.type = 255; cur.node = 0; cur.pad2 = 0; cur.thr[0] = cur.thr[1] = cur.thr[2] = 0.f; }
    if (valid) {
      load_ray_cs(&cur, rays + i);
      if (!PLASTIC && cur.type == RAY_DEAD) { /* filler of a chunk: nothing was traced */ }
      else if (!PLASTIC || cur.type != RAY_SHADOW_HEAD) { load_hit_cs(&hr, a.hits + i); cnt.rays[cur.type]++; }
    }
    if constexpr (PLASTIC) {
      const bool is_head = cur.type == RAY_SHADOW_HEAD, is_shadow = cur.type == RAY_SHADOW;
      C3 term = c3(0, 0, 0);
      if (is_shadow) {
        if (hr.inst >= 0) { cnt.hits++; cnt.levels += a.sc.meshes[a.sc.inst[hr.inst].mesh].log2_tris; }       // counted like the inline shadow rays
        term = shadow_term(a.sc, cur.thr, cur.pad2, hr);
      }
      const int n = is_head ? (int)(unsigned)(cur.node & 0xffffffffull) : 0;
      const int n_in = min(n, 31 - lane);                     // records of the block inside this warp's 32
      int maxn = n_in;
      for (int o = 16; o > 0; o >>= 1) maxn = max(maxn, __shfl_xor_sync(0xffffffffu, maxn, o));
      C3 diff = c3(0, 0, 0);
      for (int k = 1; k <= maxn; k++) {
        const float tr = __shfl_sync(0xffffffffu, term.r, (lane + k) & 31), tg = __shfl_sync(0xffffffffu, term.g, (lane + k) & 31),
                    tb = __shfl_sync(0xffffffffu, term.b, (lane + k) & 31);
        if (k <= n_in) { diff.r = fadd(diff.r, tr); diff.g = fadd(diff.g, tg); diff.b = fadd(diff.b, tb); }
      }
      if (is_head) {
        for (int k = n_in + 1; k <= n; k++) {                 // the rest of the block belongs to the next warp's records
          const RayRec *rr = rays + i + k;
          const float4 w = __ldcs(reinterpret_cast<const float4 *>(rr) + 4);          // thr[3], slot
          const int kdbits = __ldcs(reinterpret_cast<const int *>(rr) + 26);           // pad2
          HitRec h2; load_hit_cs(&h2, a.hits + i + k);
          const float th[3] = {w.x, w.y, w.z};
          const C3 t2 = shadow_term(a.sc, th, kdbits, h2);
          diff.r = fadd(diff.r, t2.r); diff.g = fadd(diff.g, t2.g); diff.b = fadd(diff.b, t2.b);
        }
        // Cs = diff * diffuse * diff_map, times the parent's throughput (plastic_shader.cc:148-156)
        const DShader &sh = a.sc.shaders[cur.filter_shader];
        const float dmx = __uint_as_float((unsigned)(cur.node >> 32)), dmy = __int_as_float(cur.pad2), dmz = __int_as_float(cur.pad3);
        const long long fr_ = to_fix(fmul(cur.thr[0], fmul(fmul(diff.r, sh.diffuse[0]), dmx)));
        const long long fg_ = to_fix(fmul(cur.thr[1], fmul(fmul(diff.g, sh.diffuse[1]), dmy)));
        const long long fb_ = to_fix(fmul(cur.thr[2], fmul(fmul(diff.b, sh.diffuse[2]), dmz)));
        Accum *acc = a.accum + cur.slot;
        if (fr_) atomicAdd((unsigned long long *)&acc->r, (unsigned long long)fr_);
        if (fg_) atomicAdd((unsigned long long *)&acc->g, (unsigned long long)fg_);
        if (fb_) atomicAdd((unsigned long long *)&acc->b, (unsigned long long)fb_);
      }
      if (is_head || is_shadow) continue;
    }
    if (!valid || hr.inst < 0) continue;
    int ti, x, y; TileGrid g;
    slot_decode(a.fr, a.tiles, a.wstride, cur.slot, &ti, &g, &x, &y);
    PathKey key; key.seed = a.fr.seed; key.tile = (uint32_t)a.tiles[ti].id; key.sample = (uint32_t)(y * g.nsx + x);
    Shading<T, PLASTIC, true> sh(a.sc, a.fr, key, cnt);
    QueueSink sink{a, a.queue[a.cur ^ 1], lane, 0, 0, 0, false, 0.f, wc};
    Hit h; h.t = hr.t; h.u = hr.u; h.v = hr.v; h.prim = hr.prim; h.inst = hr.inst;
    sh.shade(cur, h, sink);
    Accum *acc = a.accum + cur.slot;
    if (sink.r) atomicAdd((unsigned long long *)&acc->r, (unsigned long long)sink.r);
    if (sink.g) atomicAdd((unsigned long long *)&acc->g, (unsigned long long)sink.g);
    if (sink.b) atomicAdd((unsigned long long *)&acc->b, (unsigned long long)sink.b);
    if (sink.has_alpha) acc->a = sink.av;
  }
  if (wc) {                                 // fill what the warp reserved and did not use
    __syncwarp();
    RayRec dead; memset(&dead, 0, sizeof dead);
    dead.d[2] = 1.; dead.tmin = 1.; dead.tmax = 0.; dead.type = RAY_DEAD; dead.target = a.fr.target_group; dead.filter_shader = -1;
    RayRec *next = a.queue[a.cur ^ 1];
    for (unsigned p = wc->used + (unsigned)lane; p < wc->alloc; p += 32u) {
      const unsigned slot = wc->base[(p / FJ_QCHUNK) & 1u] + (p % FJ_QCHUNK);
      if (slot < a.capacity) store_ray_cs(next + slot, dead);
    }
  }
  flush_counters(cnt, 0, a.counters, lane);
}

// ------------------------------------------------------------------------------------------ resolve
// reconstruct_image + apply_pixel_filter (src/fj_renderer.cc:939-995), get_sampleset_in_pixel
// (src/fj_fixed_grid_sampler.cc:97-124), Gaussian (src/fj_filter.cc:49-58).  One CTA per tile, one thread per pixel.
// Output: packed tile blocks, block ti = bw*bh float4 (row-major inside the tile; texels outside the tile untouched).
#define FJ_RESOLVE_SPLIT 4
__global__ void __launch_bounds__(256) k_resolve_tiles(const DFrame fr, const DTile *tiles, uint32_t wstride,
                                                       const Accum *samples, float4 *blocks, int bw, int bh) {
  const int ti = blockIdx.x / FJ_RESOLVE_SPLIT, part = blockIdx.x % FJ_RESOLVE_SPLIT;      // a tile is shared by FJ_RESOLVE_SPLIT CTAs
  const DTile t = tiles[ti];
  const TileGrid g = tile_grid(fr, t);
  const int w = t.xmax - t.xmin, h = t.ymax - t.ymin;
  const int npx = fr.xrate + 2 * fr.mx, npy = fr.yrate + 2 * fr.my;
  const Accum *smp = samples + (size_t)ti * wstride;
  // `2 fx / filterwidth`: when the width is a power of two (the default 2 is) the division equals the multiplication by its
  // reciprocal bit for bit — one FP64 multiply instead of an FP64 division per axis per (pixel, sample) pair
  const bool pow2x = (__double_as_longlong(fr.xfw) & 0x000fffffffffffffll) == 0 && fr.xfw > 1e-300 && fr.xfw < 1e300;
  const bool pow2y = (__double_as_longlong(fr.yfw) & 0x000fffffffffffffll) == 0 && fr.yfw > 1e-300 && fr.yfw < 1e300;
  const double rxfw = ddiv(1., fr.xfw), ryfw = ddiv(1., fr.yfw);
  for (int p = part * blockDim.x + threadIdx.x; p < w * h; p += FJ_RESOLVE_SPLIT * blockDim.x) {
    const int px = p % w, py = p / w;
    const int x = t.xmin + px, y = t.ymin + py;
    float pr = 0, pg = 0, pb = 0, pa = 0, wsum = 0;
    for (int sy = 0; sy < npy; sy++) {
      const int gy = py * fr.yrate + sy;
      for (int sx = 0; sx < npx; sx++) {
        const int gx = px * fr.xrate + sx;
        double u, v; sample_uv(fr, g, gx, gy, &u, &v);
        const Accum s = smp[sample_slot(g, gx, gy)];
        const double fx = dsub(dmul((double)fr.xres, u), dadd((double)x, .5));
        const double fy = dsub(dmul((double)fr.yres, dsub(1., v)), dadd((double)y, .5));
        const double xx = pow2x ? dmul(dmul(2., fx), rxfw) : ddiv(dmul(2., fx), fr.xfw);
        const double yy = pow2y ? dmul(dmul(2., fy), ryfw) : ddiv(dmul(2., fy), fr.yfw);
        const double wgt = exp(dmul(-2., dadd(dmul(xx, xx), dmul(yy, yy))));
        // float accumulators, double products (fj_renderer.cc:953-961: `pixel.r += wgt * sample.data.r`)
        pr = (float)dadd((double)pr, dmul(wgt, (double)from_fix(s.r)));
        pg = (float)dadd((double)pg, dmul(wgt, (double)from_fix(s.g)));
        pb = (float)dadd((double)pb, dmul(wgt, (double)from_fix(s.b)));
        pa = (float)dadd((double)pa, dmul(wgt, (double)s.a));
        wsum = (float)dadd((double)wsum, wgt);
      }
    }
    const float inv = __fdiv_rn(1.f, wsum);
    blocks[((size_t)ti * bh + py) * bw + px] = make_float4(fmul(pr, inv), fmul(pg, inv), fmul(pb, inv), fmul(pa, inv));
  }
}

// Scatter packed tile blocks into a row-major frame (device-resident frame of the bench leg).
__global__ void k_blocks_to_frame(const DTile *tiles, int ntiles, const float4 *blocks, int bw, int bh, float4 *frame, int xres) {
  const int ti = blockIdx.x;
  const DTile t = tiles[ti];
  const int w = t.xmax - t.xmin, h = t.ymax - t.ymin;
  for (int p = threadIdx.x; p < w * h; p += blockDim.x) {
    const int px = p % w, py = p / w;
    frame[(size_t)(t.ymin + py) * xres + t.xmin + px] = blocks[((size_t)ti * bh + py) * bw + px];
  }
}

// Un-permutes the all-gathered tile blocks of R ranks into the row-major frame: tile i was rendered by rank i % R as that
// rank's block i / R; every rank contributed `per` blocks (short ranks pad), so the block sits at (i % R) * per + i / R.
__global__ void k_gathered_to_frame(const DTile *tiles, int ntiles, int nranks, int per, const float4 *gathered, int bw, int bh, float4 *frame, int xres) {
  const int ti = blockIdx.x;
  const DTile t = tiles[ti];
  const float4 *blk = gathered + ((size_t)(ti % nranks) * per + (size_t)(ti / nranks)) * bw * bh;
  const int w = t.xmax - t.xmin, h = t.ymax - t.ymin;
  for (int p = threadIdx.x; p < w * h; p += blockDim.x) {
    const int px = p % w, py = p / w;
    frame[(size_t)(t.ymin + py) * xres + t.xmin + px] = blk[(size_t)py * bw + px];
  }
}

// ------------------------------------------------------------------------------------------ probes
// Closest hit of caller-supplied rays through the megakernel's traversal (FP32 or FP64 box culling).
template <typename T>
__global__ void __launch_bounds__(128) k_trace_closest(const DScene sc, int group, int n, const double *orig, const double *dir,
                                                       const double *tmin, const double *tmax,
                                                       double *out_t, double *out_u, double *out_v, int32_t *out_prim, int32_t *out_inst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RayD r; r.o = mk(orig[3 * i], orig[3 * i + 1], orig[3 * i + 2]); r.d = mk(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
  r.tmin = tmin[i]; r.tmax = tmax[i];
  Hit h;
  const bool hit = trace_closest<T>(sc, group, r, &h);
  out_t[i] = hit ? h.t : FJ_REAL_MAX; out_u[i] = hit ? h.u : 0.; out_v[i] = hit ? h.v : 0.;
  out_prim[i] = hit ? h.prim : -1; out_inst[i] = hit ? h.inst : -1;
}
// The same probe through the wavefront's extend kernel: rays -> queue records, hit records -> arrays.
__global__ void k_probe_pack(int group, int n, const double *orig, const double *dir, const double *tmin, const double *tmax, RayRec *q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RayRec r; memset(&r, 0, sizeof r);
  for (int k = 0; k < 3; k++) { r.o[k] = orig[3 * i + k]; r.d[k] = dir[3 * i + k]; }
  r.tmin = tmin[i]; r.tmax = tmax[i]; r.target = group; r.filter_shader = -1;
  q[i] = r;
}
__global__ void k_probe_unpack(int n, const HitRec *hits, double *out_t, double *out_u, double *out_v, int32_t *out_prim, int32_t *out_inst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const HitRec h = hits[i];
  const bool hit = h.inst >= 0;
  out_t[i] = hit ? h.t : FJ_REAL_MAX; out_u[i] = hit ? h.u : 0.; out_v[i] = hit ? h.v : 0.;
  out_prim[i] = hit ? h.prim : -1; out_inst[i] = h.inst;
}

// Per-sample dump of one tile (probe): uv + radiance in row-major sample order.
__global__ void k_dump_tile_samples(const DFrame fr, const DTile t, const Accum *samples, double *out_uv, float4 *out_rgba) {
  const TileGrid g = tile_grid(fr, t);
  const int n = g.nsx * g.nsy;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int x = i % g.nsx, y = i / g.nsx;
    double u, v; sample_uv(fr, g, x, y, &u, &v);
    out_uv[2 * i] = u; out_uv[2 * i + 1] = v;
    const Accum s = samples[sample_slot(g, x, y)];
    out_rgba[i] = make_float4(from_fix(s.r), from_fix(s.g), from_fix(s.b), s.a);
  }
}

}  // namespace fj
