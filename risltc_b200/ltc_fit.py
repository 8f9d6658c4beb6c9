"""Generate plausible GGX linearly-transformed-cosine fits in the reference's
``fit<i>.dat`` layout (ltc_table.c:46-47,82-84), because the published fit data
(data/ggx_ltc_fit, README.md:8-11) is a website download that is unavailable
offline.

This is a moment-matching fitter, not Heitz's Nelder-Mead fit: for every
(roughness, inclination) cell it integrates the specular BRDF of brdfs.glsl:58-93
(GGX, height-correlated Smith) times cosine by importance sampling of the GGX
half-vector, takes the mean direction as the lobe axis and matches the second
moments of the lobe across / along the plane of incidence with those of a
scaled clamped cosine. The Fresnel layer only changes the albedo, which is
affine in F0 for Schlick Fresnel.

Record layout written per cell: (M00, M20, M11, M02, albedo) of the
cosine->shading matrix M = [[M00,0,M02],[0,M11,0],[M20,0,1]]: the loader forms the
adjugate of [[a,0,b],[0,c,0],[d,0,1]] and the shader reads it transposed
(ltc_table.c:86-90,103 -> ltc_utility.glsl:69-72), so b is M20 and d is M02.
"""
import numpy as np


def _isotropic_second_moment_table():
    """E[x^2] of normalize(s*x, s*y, z) for (x, y, z) ~ clamped cosine, as a function of s."""
    n = 256
    u = (np.arange(n) + 0.5) / n
    r2, phi = np.meshgrid(u, 2.0 * np.pi * u, indexing="ij")
    r = np.sqrt(r2)
    x, y, z = r * np.cos(phi), r * np.sin(phi), np.sqrt(1.0 - r2)
    s = np.concatenate([[0.0], np.geomspace(1e-4, 64.0, 200)])
    g = np.empty_like(s)
    for i, si in enumerate(s):
        nx, ny = si * x, si * y
        g[i] = np.mean(nx * nx / (nx * nx + ny * ny + z * z))
    return s, g


def _fit_row(res, row, samples, alpha_min, s_tab, g_tab):
    """One inclination row of the fit: returns (m00, m02, m11, m20, albedo_F1, albedo_schlick), each (res,)."""
    sqrt_alpha = np.arange(res) / (res - 1)
    alpha = np.maximum(sqrt_alpha ** 2, alpha_min)[:, None]                 # (res, 1)
    theta = min(row / (res - 1) * (0.5 * np.pi), np.radians(89.0))
    u = (np.arange(samples) + 0.5) / samples
    u1, u2 = [a.reshape(1, -1) for a in np.meshgrid(u, u, indexing="ij")]
    v = np.array([np.sin(theta), 0.0, np.cos(theta)])
    # GGX half-vector sampling: cos^2(theta_h) = (1 - u1) / (1 + (alpha^2 - 1) u1)
    a2 = alpha * alpha
    cos_h2 = (1.0 - u1) / (1.0 + (a2 - 1.0) * u1)
    cos_h = np.sqrt(cos_h2)
    sin_h = np.sqrt(np.maximum(0.0, 1.0 - cos_h2))
    phi = 2.0 * np.pi * u2
    h = np.stack([sin_h * np.cos(phi), sin_h * np.sin(phi), cos_h * np.ones_like(phi)], axis=-1)  # (res,S,3)
    v_dot_h = h @ v
    l = 2.0 * v_dot_h[..., None] * h - v
    n_l = l[..., 2]
    n_v = v[2]
    valid = (n_l > 0.0) & (v_dot_h > 0.0)
    n_lc = np.where(valid, n_l, 1.0)
    masking = n_lc * np.sqrt((-n_v * a2 + n_v) * n_v + a2)
    shadowing = n_v * np.sqrt((-n_lc * a2 + n_lc) * n_lc + a2)
    vis = 0.5 / (masking + shadowing)
    # weight = f * cos / pdf with f = D V F, pdf_l = D (n.h) / (4 v.h)
    w = np.where(valid, 4.0 * vis * n_lc * v_dot_h / np.maximum(cos_h, 1e-8), 0.0)
    schlick = np.where(valid, (1.0 - np.clip(v_dot_h, 0.0, 1.0)) ** 5, 0.0)
    norm = w.mean(axis=-1)                                  # albedo with F = 1
    norm_schlick = (w * schlick).mean(axis=-1)
    w_sum = np.maximum(w.sum(axis=-1), 1e-30)
    mean = (w[..., None] * l).sum(axis=-2) / w_sum[..., None]
    mean[..., 1] = 0.0
    mean /= np.maximum(np.linalg.norm(mean, axis=-1, keepdims=True), 1e-30)
    # never tilt the axis below 80 degrees so that M22 stays well away from zero
    phi_axis = np.clip(np.arctan2(mean[..., 0], mean[..., 2]), -np.radians(80.0), np.radians(80.0))
    cx, sx = np.cos(phi_axis), np.sin(phi_axis)
    x_axis = np.stack([cx, np.zeros_like(cx), -sx], axis=-1)
    mx = (w * np.sum(l * x_axis[:, None, :], axis=-1) ** 2).sum(axis=-1) / w_sum
    my = (w * l[..., 1] ** 2).sum(axis=-1) / w_sum
    scale_x = np.clip(np.interp(np.clip(mx, g_tab[0], g_tab[-1]), g_tab, s_tab), 1e-3, 16.0)
    scale_y = np.clip(np.interp(np.clip(my, g_tab[0], g_tab[-1]), g_tab, s_tab), 1e-3, 16.0)
    # M = R_y(phi) * diag(sx, sy, 1), divided by M22 = cos(phi)
    return scale_x, sx / cx, scale_y / cx, -scale_x * sx / cx, norm, norm_schlick


def fit_ggx_ltc(resolution=64, fresnel_count=51, samples=64, alpha_min=2.0e-3):
    """Return fits of shape (fresnel_count, resolution, resolution, 5), float32.

    Axis 1 is inclination (row, theta = row/(res-1) * pi/2), axis 2 is roughness
    (column, sqrt(alpha) = col/(res-1)), matching ltc_table.h:47-51 and the lookup
    coordinates of ltc_utility.glsl:63-66."""
    res = resolution
    s_tab, g_tab = _isotropic_second_moment_table()
    rows = [_fit_row(res, row, samples, alpha_min, s_tab, g_tab) for row in range(res)]
    m00, m02, m11, m20, norm, norm_schlick = [np.stack([r[i] for r in rows]) for i in range(6)]
    fits = np.empty((fresnel_count, res, res, 5), dtype=np.float32)
    for k in range(fresnel_count):
        f0 = k / max(fresnel_count - 1, 1)
        albedo = np.clip(f0 * (norm - norm_schlick) + norm_schlick, 0.0, 1.0)
        fits[k, ..., 0] = m00
        fits[k, ..., 1] = m20
        fits[k, ..., 2] = m11
        fits[k, ..., 3] = m02
        fits[k, ..., 4] = albedo
    return fits


def quantize_fits(fits):
    """numpy restatement of the loader's quantisation (ltc_table.c:82-116) for host tooling.

    Returns (rgba16 (layers,res,res,4), rg16 (layers,res,res,2)) as uint16."""
    d = np.asarray(fits, dtype=np.float32)
    a, b, c, dd, albedo = [d[..., i] for i in range(5)]
    z = np.zeros_like(a)
    inv = np.stack([np.stack([c, z, -b * c], -1), np.stack([z, a - b * dd, z], -1), np.stack([-c * dd, z, a * c], -1)], -2)
    mag = np.abs(inv).reshape(inv.shape[:-2] + (9,)).max(axis=-1)
    inv = (inv / mag[..., None, None]).astype(np.float32)
    proc = np.stack([inv[..., 0, 0], -inv[..., 0, 2], inv[..., 1, 1], inv[..., 2, 0], inv[..., 2, 2], albedo], -1)
    proc = np.clip(proc, 0.0, 1.0).astype(np.float32)
    q = (proc * np.float32(65535.0) + np.float32(0.5)).astype(np.uint16)
    return np.ascontiguousarray(q[..., :4]), np.ascontiguousarray(q[..., 4:])
