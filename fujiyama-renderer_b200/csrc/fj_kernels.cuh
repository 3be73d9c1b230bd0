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
// Camera::GetRay, src/fj_camera.cc:79-110 (static camera: one matrix)
__device__ __forceinline__ void camera_ray(const DCamera &c, double u, double v, RayD *ray) {
  const D3 target = mat_point(c.fwd, mk(dmul(dsub(u, .5), c.uvx), dmul(dsub(v, .5), c.uvy), -1.));
  const D3 eye = mat_point(c.fwd, mk(0., 0., 0.));
  ray->d = normalize(target - eye); ray->o = eye; ray->tmin = c.znear; ray->tmax = c.zfar;
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

// ------------------------------------------------------------------------------------------ path
struct PathKey { uint32_t seed, tile, sample; };

template <typename T>
struct PathTracer {
  const DScene &sc; const DFrame &fr; PathKey key;
  unsigned long long rays[5];
  unsigned int hits, levels;
  C3 acc;
  Pending stack[FJ_PENDING];
  int sp;

  __device__ PathTracer(const DScene &s, const DFrame &f, PathKey k) : sc(s), fr(f), key(k), sp(0) {
    rays[0] = rays[1] = rays[2] = rays[3] = rays[4] = 0; hits = levels = 0; acc = c3(0, 0, 0);
  }
  __device__ __forceinline__ double rnd(unsigned long long node, uint32_t dim) const { return ctr_rand(key.seed, key.tile, key.sample, node, dim); }

  // SlIlluminance (src/fj_shading.cc:296-359) for one light sample; returns Kd * Cl contribution inputs.
  __device__ __forceinline__ bool illuminance(const DLight &lt, const LightSample &ls, const D3 &Ps, const D3 &axis,
                                              int shaded_object, C3 *Cl, D3 *Ln_out) {
    D3 Ln = ls.P - Ps;
    const double distance = length(Ln);
    if (distance > 0) { const double inv = ddiv(1., distance); Ln = Ln * inv; }
    *Ln_out = Ln;
    const D3 nml_axis = normalize(axis);
    const double cosangle = dot(nml_axis, Ln);
    if (cosangle < 6.123233995736766e-17) return false;          // cos(PI/2) in FP64
    C3 lc = light_illuminate(lt, ls, Ps);
    if ((double)lc.r < .0001 && (double)lc.g < .0001 && (double)lc.b < .0001) return false;
    if (fr.cast_shadow) {                                        // SlShadowContext :266-279
      RayD sr; sr.o = Ps; sr.d = Ln; sr.tmin = .0001; sr.tmax = distance;
      Hit h;
      rays[RAY_SHADOW]++;
      if (trace_closest<T>(sc, sc.inst[shaded_object].shadow_target, sr, &h)) {
        hits++; levels += sc.meshes[sc.inst[h.inst].mesh].log2_tris;
        const float ac = fadd(1.f, -occluder_opacity(sc, h));
        lc.r = fmul(lc.r, ac); lc.g = fmul(lc.g, ac); lc.b = fmul(lc.b, ac);
      }
    }
    *Cl = lc;
    return true;
  }

  // diff = sum over all light samples of max(0, Nf.Ln) * Cl — the loop of PlasticShader::evaluate
  // (plastic_shader.cc:117-137) over SlNewLightSamples (fj_shading.cc:380-404; Light::GetSamples).
  __device__ C3 gather_lights(const D3 &P, const D3 &Nf, int shaded_object, unsigned long long node) {
    C3 diff = c3(0, 0, 0);
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
          const double x = dsub(rnd(node, dim), .5), z = dsub(rnd(node, dim + 1), .5); dim += 2;
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
        C3 Cl; D3 Ln;
        if (!illuminance(lt, ls, P, Nf, shaded_object, &Cl, &Ln)) continue;
        float Kd = (float)dot(Nf, Ln);
        Kd = Kd > 0.f ? Kd : 0.f;
        diff.r = fadd(diff.r, fmul(Kd, Cl.r)); diff.g = fadd(diff.g, fmul(Kd, Cl.g)); diff.b = fadd(diff.b, fmul(Kd, Cl.b));
      }
    }
    return diff;
  }

  __device__ __forceinline__ void add(const C3 &thr, float r, float g, float b) {
    acc.r = fadd(acc.r, fmul(thr.r, r)); acc.g = fadd(acc.g, fmul(thr.g, g)); acc.b = fadd(acc.b, fmul(thr.b, b));
  }
  __device__ __forceinline__ void push(const Pending &p) { if (sp < FJ_PENDING) stack[sp++] = p; }

  // SlTrace for a camera sample and everything it spawns.  Returns (rgb, alpha) of the sample.
  __device__ float4 run(const RayD &cam) {
    float alpha = 0.f;
    Pending cur;
    cur.o = cam.o; cur.d = cam.d; cur.tmin = cam.tmin; cur.thr = c3(1, 1, 1); cur.transmit = c3(1, 1, 1);
    cur.node = 1; cur.target = fr.target_group; cur.type = RAY_CAMERA; cur.dd = cur.rd = cur.fd = 0; cur.filter = 0;
    double tmax = cam.tmax;
    bool have = true;
    while (have || sp > 0) {
      if (!have) { cur = stack[--sp]; tmax = 1000.; }
      have = false;
      RayD ray; ray.o = cur.o; ray.d = cur.d; ray.tmin = cur.tmin; ray.tmax = tmax;
      rays[cur.type]++;
      Hit h;
      if (!trace_closest<T>(sc, cur.target, ray, &h)) continue;
      hits++; levels += sc.meshes[sc.inst[h.inst].mesh].log2_tris;
      C3 thr = cur.thr;
      if (cur.filter) {    // pathtracing_shader.cc:247-251: C *= pow(transmit, t_hit) of the refracted child
        thr.r = fmul(thr.r, (float)pow((double)cur.transmit.r, h.t));
        thr.g = fmul(thr.g, (float)pow((double)cur.transmit.g, h.t));
        thr.b = fmul(thr.b, (float)pow((double)cur.transmit.b, h.t));
      }
      D3 P, N; int slot;
      hit_surface(sc, ray, h, &P, &N, &slot);
      const DInstance &in = sc.inst[h.inst];
      float Os = 1.f;
      const int kind = slot < 0 ? 0 : sc.shaders[slot].kind;
      if (kind == 0) {                                             // NO_SHADER_COLOR, fj_shading.cc:26,555-560
        add(thr, .5f, 1.f, 0.f);
      } else if (kind == 1) {                                      // ConstantShader::evaluate, constant_shader.cc:72-94
        const DShader &sh = sc.shaders[slot];
        add(thr, sh.diffuse[0], sh.diffuse[1], sh.diffuse[2]);
      } else if (kind == 2) {                                      // PlasticShader::evaluate, plastic_shader.cc:101-179
        const DShader &sh = sc.shaders[slot];
        const D3 Nf = sl_faceforward(ray.d, N);
        const C3 diff = gather_lights(P, Nf, h.inst, cur.node);
        add(thr, fmul(diff.r, sh.diffuse[0]), fmul(diff.g, sh.diffuse[1]), fmul(diff.b, sh.diffuse[2]));
        if (sh.do_reflect && (int)cur.rd + 1 <= fr.max_reflect) {  // SlReflectContext :242-252, gate :467-499
          const double Kr = sl_fresnel(ray.d, Nf, ddiv(1., (double)sh.ior));
          Pending c;
          c.o = P; c.d = normalize(sl_reflect(ray.d, Nf)); c.tmin = .001;
          c.thr = c3(fmul(thr.r, (float)dmul(Kr, (double)sh.reflect[0])), fmul(thr.g, (float)dmul(Kr, (double)sh.reflect[1])),
                     fmul(thr.b, (float)dmul(Kr, (double)sh.reflect[2])));
          c.transmit = c3(1, 1, 1); c.filter = 0;
          c.node = cur.node * 4 + 2; c.target = in.reflect_target; c.type = RAY_REFLECT;
          c.dd = cur.dd; c.rd = cur.rd + 1; c.fd = cur.fd;
          push(c);
        }
        Os = sh.opacity;
      } else {                                                     // PathtracingShader::evaluate, pathtracing_shader.cc:125-257
        const DShader &sh = sc.shaders[slot];
        add(thr, sh.emission[0], sh.emission[1], sh.emission[2]);
        if (luminance(sh.refract) > 0.f && (int)cur.fd + 1 <= fr.max_refract) {     // integrate_refract :231-257
          const double ior = ddiv(1., (double)sh.ior);
          const double Kt = dsub(1., sl_fresnel(ray.d, N, ior));
          Pending c;
          c.o = P; c.d = normalize(sl_refract(ray.d, N, ior)); c.tmin = .0001;
          const float kt = (float)Kt;
          c.thr = c3(fmul(thr.r, fmul(kt, sh.refract[0])), fmul(thr.g, fmul(kt, sh.refract[1])), fmul(thr.b, fmul(kt, sh.refract[2])));
          c.filter = (sh.do_color_filter && dot(ray.d, N) < 0) ? 1 : 0;
          c.transmit = c3(sh.transmit[0], sh.transmit[1], sh.transmit[2]);
          c.node = cur.node * 4 + 3; c.target = in.refract_target; c.type = RAY_REFRACT;
          c.dd = cur.dd; c.rd = cur.rd; c.fd = cur.fd + 1;
          push(c);
        }
        if (luminance(sh.reflect) > 0.f && (int)cur.rd + 1 <= fr.max_reflect) {     // integrate_reflect :210-229
          const double Kr = sl_fresnel(ray.d, N, ddiv(1., (double)sh.ior));
          Pending c;
          c.o = P; c.d = normalize(sl_reflect(ray.d, N)); c.tmin = .001;
          const float kr = (float)Kr;
          c.thr = c3(fmul(thr.r, fmul(kr, sh.reflect[0])), fmul(thr.g, fmul(kr, sh.reflect[1])), fmul(thr.b, fmul(kr, sh.reflect[2])));
          c.transmit = c3(1, 1, 1); c.filter = 0;
          c.node = cur.node * 4 + 2; c.target = in.reflect_target; c.type = RAY_REFLECT;
          c.dd = cur.dd; c.rd = cur.rd + 1; c.fd = cur.fd;
          push(c);
        }
        if (luminance(sh.diffuse) > 0.f && (int)cur.dd + 1 <= fr.max_diffuse) {     // integrate_diffuse :176-208
          const D3 w = N;
          D3 u = fabs(w.x) > .001 ? mk(0., 1., 0.) : mk(1., 0., 0.);
          u = normalize(cross(u, w));
          const D3 v = cross(w, u);
          const double x1 = rnd(cur.node, 0), x2 = rnd(cur.node, 1);
          const double r1 = dmul(dmul(2., 3.14159265358979323846), x1), r2 = x2, r2s = __dsqrt_rn(r2);
          const D3 D = normalize(((u * cos(r1)) * r2s + (v * sin(r1)) * r2s) + w * __dsqrt_rn(dsub(1., r2)));
          const float Kd = (float)dot(N, D);
          // the next ray continues in registers (it is the deepest branch of the DFS)
          cur.thr = c3(fmul(thr.r, fmul(Kd, sh.diffuse[0])), fmul(thr.g, fmul(Kd, sh.diffuse[1])), fmul(thr.b, fmul(Kd, sh.diffuse[2])));
          cur.o = P; cur.d = D; cur.tmin = .001; cur.transmit = c3(1, 1, 1); cur.filter = 0;
          cur.node = cur.node * 4 + 1; cur.target = in.reflect_target;
          const bool was_camera = cur.type == RAY_CAMERA;
          cur.type = RAY_DIFFUSE; cur.dd = cur.dd + 1;
          tmax = 1000.; have = true;
          if (was_camera) alpha = 1.f;
          continue;
        }
      }
      if (cur.type == RAY_CAMERA) alpha = fminf(fmaxf(Os, 0.f), 1.f);             // trace_surface :562-566
    }
    return make_float4(acc.r, acc.g, acc.b, alpha);
  }
};

// ------------------------------------------------------------------------------------------ kernels
struct RenderArgs {
  DScene sc; DCamera cam; DFrame fr;
  const DTile *tiles; int ntiles;         // tiles of this batch
  uint32_t wstride;                       // sample slots per tile in `samples` (multiple of 32)
  float4 *samples;                        // ntiles * wstride
  DCounters *counters;
  unsigned long long *work;               // global work counter (units of 32 slots)
};

template <typename T>
__global__ void __launch_bounds__(128) k_render_samples(const RenderArgs a) {
  const int lane = threadIdx.x & 31;
  const unsigned long long total = (unsigned long long)a.ntiles * a.wstride;
  unsigned long long cnt[7] = {0, 0, 0, 0, 0, 0, 0}; unsigned long long nsamp = 0;
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
    float4 out = make_float4(0, 0, 0, 0);
    if (x < g.nsx && y < g.nsy) {
      double u, v; sample_uv(a.fr, g, x, y, &u, &v);
      RayD ray; camera_ray(a.cam, u, v, &ray);
      PathKey key; key.seed = a.fr.seed; key.tile = (uint32_t)tile.id; key.sample = (uint32_t)(y * g.nsx + x);
      PathTracer<T> pt(a.sc, a.fr, key);
      out = pt.run(ray);
      for (int k = 0; k < 5; k++) cnt[k] += pt.rays[k];
      cnt[5] += pt.hits; cnt[6] += pt.levels;
      nsamp++;
    }
    a.samples[(size_t)ti * a.wstride + (blk << 5) + lane] = out;
  }
  for (int k = 0; k < 7; k++) {
    unsigned long long v = cnt[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0 && v) atomicAdd(k < 5 ? &a.counters->rays[k] : (k == 5 ? &a.counters->hits : &a.counters->levels), v);
  }
  for (int o = 16; o > 0; o >>= 1) nsamp += __shfl_down_sync(0xffffffffu, nsamp, o);
  if (lane == 0 && nsamp) atomicAdd(&a.counters->samples, nsamp);
}

// reconstruct_image + apply_pixel_filter (src/fj_renderer.cc:939-995), get_sampleset_in_pixel
// (src/fj_fixed_grid_sampler.cc:97-124), Gaussian (src/fj_filter.cc:49-58).  One CTA per tile, one thread per pixel.
// Output: packed tile blocks, block ti = bw*bh float4 (row-major inside the tile; texels outside the tile untouched).
__global__ void __launch_bounds__(256) k_resolve_tiles(const DFrame fr, const DTile *tiles, uint32_t wstride,
                                                       const float4 *samples, float4 *blocks, int bw, int bh) {
  const int ti = blockIdx.x;
  const DTile t = tiles[ti];
  const TileGrid g = tile_grid(fr, t);
  const int w = t.xmax - t.xmin, h = t.ymax - t.ymin;
  const int npx = fr.xrate + 2 * fr.mx, npy = fr.yrate + 2 * fr.my;
  const float4 *smp = samples + (size_t)ti * wstride;
  for (int p = threadIdx.x; p < w * h; p += blockDim.x) {
    const int px = p % w, py = p / w;
    const int x = t.xmin + px, y = t.ymin + py;
    float pr = 0, pg = 0, pb = 0, pa = 0, wsum = 0;
    for (int sy = 0; sy < npy; sy++) {
      const int gy = py * fr.yrate + sy;
      for (int sx = 0; sx < npx; sx++) {
        const int gx = px * fr.xrate + sx;
        double u, v; sample_uv(fr, g, gx, gy, &u, &v);
        const float4 s = smp[sample_slot(g, gx, gy)];
        const double fx = dsub(dmul((double)fr.xres, u), dadd((double)x, .5));
        const double fy = dsub(dmul((double)fr.yres, dsub(1., v)), dadd((double)y, .5));
        const double xx = ddiv(dmul(2., fx), fr.xfw), yy = ddiv(dmul(2., fy), fr.yfw);
        const double wgt = exp(dmul(-2., dadd(dmul(xx, xx), dmul(yy, yy))));
        // float accumulators, double products (fj_renderer.cc:953-961: `pixel.r += wgt * sample.data.r`)
        pr = (float)dadd((double)pr, dmul(wgt, (double)s.x));
        pg = (float)dadd((double)pg, dmul(wgt, (double)s.y));
        pb = (float)dadd((double)pb, dmul(wgt, (double)s.z));
        pa = (float)dadd((double)pa, dmul(wgt, (double)s.w));
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

// Per-sample dump of one tile (probe): uv + radiance in row-major sample order.
__global__ void k_dump_tile_samples(const DFrame fr, const DTile t, const float4 *samples, double *out_uv, float4 *out_rgba) {
  const TileGrid g = tile_grid(fr, t);
  const int n = g.nsx * g.nsy;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int x = i % g.nsx, y = i / g.nsx;
    double u, v; sample_uv(fr, g, x, y, &u, &v);
    out_uv[2 * i] = u; out_uv[2 * i + 1] = v;
    out_rgba[i] = samples[sample_slot(g, x, y)];
  }
}

}  // namespace fj
