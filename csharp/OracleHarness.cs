// OracleHarness.cs -- headless C# restatement of the hot path (Common.hlsl, Raytrace2D.compute Trace +
// ProcessHits, AudioConvolve.compute) for machines that have dotnet.
//
// The north star asks for a headless C# reference harness run under dotnet.  No C# toolchain exists in the
// build image, so this file is NOT compiled or run here; the operative oracle is its plain-C twin
// oracle/rar_oracle.c, which obeys the same arithmetic contract (DESIGN.md section 2): binary32 throughout, a
// fused multiply-add exactly where MathF.FusedMultiplyAdd is written, the fixed sin/cos/asin kernels below,
// Q23.40 fixed-point deposits.  On .NET Core 3.0+ (x64, SSE/FMA) the two are expected to agree bit for bit.
//
//   dotnet new console -o harness && cp OracleHarness.cs harness/Program.cs && dotnet run -c Release --project harness
//
// Expected output for the bundled SmollRoom scene, frame 1 (tests/golden/smoll_frame1.npz):
//   ray_bounces 75185 nearest_tests 1503700 shadow_tests 1069199 direct_hits 699 nee_hits 11635
//   nonzero bins 5256, sum of Q23.40 deposits 870330858820
using System;
using System.Collections.Generic;

public struct Wall
{
    public float ax, ay, bx, by, nx, ny, absorption, scattering, transmission, ior;
}

public sealed class TraceParams
{
    public float sourceX, sourceY, listenerX, listenerY;
    public float listenerRadius = 0.5f, speedOfSound = 343f, inputGain = 1f;
    public int maxBounceCount = 5;
    public uint rngStateOffset = 1;
    public int rayCount = 1000, sampleRate = 48000, impulseLength = 96000;
}

public sealed class Counters
{
    public long rayBounces, nearestTests, shadowTests, directHits, neeHits;
}

public static class Oracle
{
    const float Eps = 1e-4f, Inf = 1e8f, Pi = 3.14159265f;   // Common.hlsl:4-6

    static float Fma(float a, float b, float c) => MathF.FusedMultiplyAdd(a, b, c);
    static float Dot(float ax, float ay, float bx, float by) => Fma(ax, bx, ay * by);

    // Common.hlsl:8-12: 4294967295.0 rounds to 2^32 in binary32; (float)uint rounds to nearest.
    public static float Random(ref uint state)
    {
        state = unchecked(state * 747796405u + 2891336453u);
        uint res = unchecked(((state >> (int)((state >> 28) + 4u)) ^ state) * 277803737u);
        uint v = (res >> 22) ^ res;
        return (float)v / 4294967296.0f;
    }

    // Common.hlsl:14-21
    public static float Intersect(float ox, float oy, float dx, float dy, float ax, float ay, float bx, float by)
    {
        float v1x = ox - ax, v1y = oy - ay, v2x = bx - ax, v2y = by - ay, v3x = -dy, v3y = dx;
        float dotP = Dot(v2x, v2y, v3x, v3y);
        if (MathF.Abs(dotP) < Eps) return Inf;
        float t1 = Fma(v2x, v1y, -(v2y * v1x)) / dotP;
        float t2 = Dot(v1x, v1y, v3x, v3y) / dotP;
        return (t1 >= Eps && t2 >= 0f && t2 <= 1f) ? t1 : Inf;
    }

    // Common.hlsl:23-36
    public static float IntersectCircle(float px, float py, float dx, float dy, float cx, float cy, float radius)
    {
        float lx = cx - px, ly = cy - py;
        float tca = Dot(lx, ly, dx, dy);
        if (tca < 0f) return Inf;
        float d2 = Fma(-tca, tca, Dot(lx, ly, lx, ly));
        float r2 = radius * radius;
        if (d2 > r2) return Inf;
        float thc = MathF.Sqrt(r2 - d2);
        float t0 = tca - thc, t1 = tca + thc;
        if (t0 > Eps) return t0;
        if (t1 > Eps) return t1;
        return Inf;
    }

    // Common.hlsl:38-43 in 2-D
    static bool Refract(float ix, float iy, float nx, float ny, float eta, out float tx, out float ty)
    {
        float cosi = Dot(-ix, -iy, nx, ny);
        float cost2 = 1f - (eta * eta) * (1f - cosi * cosi);
        float k = eta * cosi - MathF.Sqrt(MathF.Abs(cost2));
        float rx = Fma(k, nx, eta * ix), ry = Fma(k, ny, eta * iy);
        if (cost2 > 0f) { tx = rx; ty = ry; return true; }
        tx = 0f; ty = 0f; return false;
    }

    // fixed transcendental kernels of the arithmetic contract
    public static void SinCos(float x, out float sn, out float cs)
    {
        const float magic = 12582912f;
        float kf = Fma(x, 0.636619772f, magic) - magic;
        int q = (int)kf;
        float r = Fma(-kf, 1.5703125f, x);
        r = Fma(-kf, 4.837512969970703125e-4f, r);
        r = Fma(-kf, 7.54978995489188e-8f, r);
        float z = r * r;
        float ps = Fma(z, -1.9515295891e-4f, 8.3321608736e-3f);
        ps = Fma(z, ps, -1.6666654611e-1f);
        float s = Fma(r * z, ps, r);
        float pc = Fma(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
        pc = Fma(z, pc, 4.166664568298827e-2f);
        float c = Fma(z * z, pc, Fma(z, -0.5f, 1f));
        switch (q & 3)
        {
            case 0: sn = s; cs = c; break;
            case 1: sn = c; cs = -s; break;
            case 2: sn = -s; cs = -c; break;
            default: sn = -c; cs = s; break;
        }
    }

    public static float Asin(float x)
    {
        float a = MathF.Abs(x);
        if (a > 1f) a = 1f;
        bool big = a > 0.5f;
        float z, w;
        if (big) { z = 0.5f * (1f - a); w = MathF.Sqrt(z); } else { w = a; z = a * a; }
        float p = Fma(z, 4.2163199048e-2f, 2.4181311049e-2f);
        p = Fma(z, p, 4.5470025998e-2f);
        p = Fma(z, p, 7.4953002686e-2f);
        p = Fma(z, p, 1.6666752422e-1f);
        float r = Fma(w * z, p, w);
        if (big) r = 1.5707963267948966f - (r + r);
        return x < 0f ? -r : r;
    }

    public static long Quantize(float e)
    {
        if (float.IsNaN(e)) return 0;
        if (e > 4194304f) e = 4194304f;
        if (e < -4194304f) e = -4194304f;
        return (long)(e * 1099511627776f);   // 2^40: exact scaling, truncation toward zero
    }

    // Raytrace2D.compute:157-165 ProcessHits, as fixed point
    static void Deposit(long[] hist, TraceParams p, float t, float e)
    {
        float ts = t * (float)p.sampleRate;
        if (!(ts > -1f && ts < (float)p.impulseLength)) return;
        int idx = (int)ts;
        if (idx >= 0 && idx < p.impulseLength) hist[idx] += Quantize(e);
    }

    // Raytrace2D.compute:40-47; vector / scalar := vector * (1 / scalar)
    static bool CheckVis(Wall[] walls, float sx, float sy, float ex, float ey, float dist, Counters c)
    {
        float inv = 1f / dist;
        float dx = (ex - sx) * inv, dy = (ey - sy) * inv, lim = dist - 0.1f;
        foreach (Wall w in walls)
        {
            c.shadowTests++;
            if (Intersect(sx, sy, dx, dy, w.ax, w.ay, w.bx, w.by) < lim) return false;
        }
        return true;
    }

    // Raytrace2D.compute:49-156, one thread
    static void TraceOne(Wall[] walls, TraceParams p, uint id, long[] hist, Counters c)
    {
        uint rng = unchecked(id + p.rngStateOffset * 719393u);
        float angle = (((float)id + Random(ref rng)) / (float)p.rayCount) * 2f * Pi;
        SinCos(angle, out float diry, out float dirx);
        float posx = p.sourceX, posy = p.sourceY, energy = p.inputGain, time = 0f, dist = 0f, speed = p.speedOfSound;
        int wallDepth = 0;
        for (int i = 0; i < p.maxBounceCount; i++)
        {
            c.rayBounces++;
            float closest = Inf; int hit = -1;
            for (int w = 0; w < walls.Length; w++)
            {
                float d = Intersect(posx, posy, dirx, diry, walls[w].ax, walls[w].ay, walls[w].bx, walls[w].by);
                if (d < closest) { closest = d; hit = w; }
            }
            c.nearestTests += walls.Length;
            if (wallDepth == 0)
            {
                float dl = IntersectCircle(posx, posy, dirx, diry, p.listenerX, p.listenerY, p.listenerRadius);
                if (dl < closest && dl < Inf)
                {
                    float total = dist + dl;
                    c.directHits++;
                    Deposit(hist, p, time + dl / speed, energy / MathF.Max(1f, total * total));
                }
            }
            if (hit < 0) break;
            posx = Fma(dirx, closest, posx); posy = Fma(diry, closest, posy);
            time += closest / speed; dist += closest;
            Wall wall = walls[hit];
            float keep = 1f - wall.absorption;
            if (wallDepth == 0)
            {
                float tlx = p.listenerX - posx, tly = p.listenerY - posy;
                float dl = MathF.Sqrt(Dot(tlx, tly, tlx, tly));
                if (CheckVis(walls, Fma(wall.nx, Eps, posx), Fma(wall.ny, Eps, posy), p.listenerX, p.listenerY, dl, c))
                {
                    bool flip = Dot(dirx, diry, wall.nx, wall.ny) > 0f;
                    float enx = flip ? -wall.nx : wall.nx, eny = flip ? -wall.ny : wall.ny;
                    float invDl = 1f / dl;
                    float cosT = MathF.Max(0f, Dot(enx, eny, tlx * invDl, tly * invDl));
                    float total = dist + dl;
                    float contrib = ((energy * keep) * (cosT * 0.5f)) * (1f / (total * total));
                    if (contrib > 1e-5f) { c.neeHits++; Deposit(hist, p, time + dl / p.speedOfSound, contrib); }
                }
            }
            energy *= keep;
            if (energy < 1e-3f) break;
            bool entering = Dot(dirx, diry, wall.nx, wall.ny) < 0f;
            float nx = entering ? wall.nx : -wall.nx, ny = entering ? wall.ny : -wall.ny;
            float wallSpeed = p.speedOfSound / wall.ior;
            float nextSpeed = entering ? wallSpeed : (wallDepth <= 1 ? p.speedOfSound : wallSpeed);
            float eta = nextSpeed / speed;
            float rngVal = Random(ref rng);
            if (rngVal < wall.transmission)
            {
                Refract(dirx, diry, nx, ny, eta, out float rx, out float ry);
                if (MathF.Sqrt(Dot(rx, ry, rx, ry)) > 0f)
                {
                    if (wall.scattering > 0f)
                    {
                        float jitter = (Random(ref rng) - 0.5f) * 2f * wall.scattering;
                        SinCos(jitter, out float sj, out float cj);
                        float jx = Fma(rx, cj, -(ry * sj)), jy = Fma(rx, sj, ry * cj);
                        rx = jx; ry = jy;
                    }
                    float invLen = 1f / MathF.Sqrt(Dot(rx, ry, rx, ry));
                    dirx = rx * invLen; diry = ry * invLen;
                    speed = nextSpeed;
                    wallDepth = entering ? wallDepth + 1 : Math.Max(0, wallDepth - 1);
                    posx = Fma(dirx, Eps, posx); posy = Fma(diry, Eps, posy);
                    continue;
                }
            }
            float k2 = 2f * Dot(dirx, diry, nx, ny);
            float spx = Fma(-k2, nx, dirx), spy = Fma(-k2, ny, diry);
            float u = Fma(2f, Random(ref rng), -1f);
            SinCos(Asin(u), out float s, out float cc);
            float dfx = Fma(nx, cc, -(ny * s)), dfy = Fma(nx, s, ny * cc);
            float mx = Fma(wall.scattering, dfx - spx, spx), my = Fma(wall.scattering, dfy - spy, spy);
            float inv = 1f / MathF.Sqrt(Dot(mx, my, mx, my));
            dirx = mx * inv; diry = my * inv;
            posx = Fma(nx, Eps, posx); posy = Fma(ny, Eps, posy);
        }
    }

    // Dispatch of Trace: ceil(rayCount/64) groups of 64 threads, no bounds guard
    public static Counters Trace(Wall[] walls, TraceParams p, long[] hist)
    {
        var c = new Counters();
        long threads = ((long)p.rayCount + 63) / 64 * 64;
        for (long id = 0; id < threads; id++) TraceOne(walls, p, (uint)id, hist, c);
        return c;
    }

    // AudioConvolve.compute:13-31 on the float view of the histogram
    public static float[] Convolve(float[] input, float[] ir, int accumCount)
    {
        int outLen = input.Length + ir.Length;
        var output = new float[outLen];
        for (int n = 0; n < outLen; n++)
        {
            float sum = 0f;
            int startK = Math.Max(0, n - ir.Length + 1), endK = Math.Min(n, input.Length - 1);
            for (int k = startK; k <= endK; k++)
            {
                float v = input[k];
                if (MathF.Abs(v) > Eps) sum += v * ir[n - k];
            }
            output[n] = accumCount > 0 ? sum / accumCount : 0f;
        }
        return output;
    }

    // Helpers/SceneHelper.cs:78-98 for a unit box under (position, quaternion z/w, scale)
    public static void AddBox(List<Wall> walls, float px, float py, float qz, float qw, float sx, float sy,
                              float absorption, float scattering, float transmission, float ior)
    {
        float r00 = 1f - 2f * (qz * qz), r01 = -(2f * (qz * qw)), r10 = 2f * (qz * qw), r11 = r00;
        float winding = sx * sy >= 0f ? 1f : -1f;
        float[] lx = { -0.5f, 0.5f, 0.5f, -0.5f }, ly = { -0.5f, -0.5f, 0.5f, 0.5f };
        for (int i = 0; i < 4; i++)
        {
            int j = (i + 1) % 4;
            float x1 = lx[i] * sx, y1 = ly[i] * sy, x2 = lx[j] * sx, y2 = ly[j] * sy;
            var w = new Wall();
            w.ax = (r00 * x1 + r01 * y1) + px; w.ay = (r10 * x1 + r11 * y1) + py;
            w.bx = (r00 * x2 + r01 * y2) + px; w.by = (r10 * x2 + r11 * y2) + py;
            float dx = w.bx - w.ax, dy = w.by - w.ay;
            float len = MathF.Sqrt(dx * dx + dy * dy);
            if (len > 1e-5f) { dx /= len; dy /= len; } else { dx = 0f; dy = 0f; }
            w.nx = dy * winding; w.ny = -dx * winding;
            w.absorption = absorption; w.scattering = scattering; w.transmission = transmission; w.ior = ior;
            walls.Add(w);
        }
    }
}

public static class Program
{
    public static void Main()
    {
        // Assets/Scenes/SmollRoom.unity (SURVEY.md Appendix B)
        var walls = new List<Wall>();
        Oracle.AddBox(walls, 0f, 10f, 0f, 1f, 100f, 1f, 0.507f, 0.5f, 0.271f, 0.01f);
        Oracle.AddBox(walls, 0.01f, -5f, 0f, 1f, 100f, 1f, 0.507f, 0.5f, 0.271f, 0.01f);
        Oracle.AddBox(walls, -20f, 0f, 0.7071068f, 0.7071068f, 20f, 1f, 0.507f, 0.5f, 0.271f, 0.01f);
        Oracle.AddBox(walls, 20f, 0f, 0.7071068f, 0.7071068f, 20f, 1f, 0.507f, 0.5f, 0.271f, 0.01f);
        Oracle.AddBox(walls, -11.8f, 7.18f, 0.47792548f, 0.8784004f, 100f, 1f, 0.148f, 1f, 1f, 0.6f);
        var p = new TraceParams { sourceX = -18f, sourceY = 9f, listenerX = 0f, listenerY = -3.68f, rayCount = 15000,
                                  maxBounceCount = 5, rngStateOffset = 1, impulseLength = 72000 };
        var hist = new long[p.impulseLength];
        Counters c = Oracle.Trace(walls.ToArray(), p, hist);
        long sum = 0; int nz = 0;
        foreach (long q in hist) { sum += q; if (q != 0) nz++; }
        Console.WriteLine($"ray_bounces {c.rayBounces} nearest_tests {c.nearestTests} shadow_tests {c.shadowTests} " +
                          $"direct_hits {c.directHits} nee_hits {c.neeHits}");
        Console.WriteLine($"nonzero bins {nz}, sum of Q23.40 deposits {sum}");
        Console.WriteLine("expected: ray_bounces 75185 nearest_tests 1503700 shadow_tests 1069199 direct_hits 699 nee_hits 11635");
        Console.WriteLine("expected: nonzero bins 5256, sum of Q23.40 deposits 870330858820");
    }
}
