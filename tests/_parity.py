"""Helpers shared by the GPU parity tests (tests only; nothing here is on the product path)."""
import torch

from oracle import qwen3_tts_oracle as O

M64 = (1 << 64) - 1


def hash_uniform(seed: int, step: int, stream: int, b: int = 0) -> float:
    """Bit-for-bit the counter-based generator of csrc/sampler.cuh (hash_uniform): the device draws u = f(seed, frame,
    code group); the oracle is fed the same numbers (stochastic parity = identical uniforms)."""
    inner = (step * 1315423911 + stream * 2654435761 + b * 97 + 1) & 0xFFFFFFFF      # 32-bit unsigned arithmetic in C
    z = (seed + 0x9E3779B97F4A7C15 * inner) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z = z ^ (z >> 31)
    return float(z >> 40) * (1.0 / 16777216.0)


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def snr_db(x: torch.Tensor, ref: torch.Tensor) -> float:
    return float(10 * torch.log10(ref.double().pow(2).sum() / (x.double() - ref.double()).pow(2).sum().clamp_min(1e-30)))


def draw_boundary_distance(scores: torch.Tensor, sp: "O.SamplingParams", u: float, a: int, b: int) -> float:
    """How far (in probability mass) the uniform `u` is from the CDF boundaries that separate ids a and b in the oracle's
    own categorical distribution; a stochastic draw may only differ between two implementations when this is tiny."""
    p = torch.softmax(scores.double(), -1)
    cdf = p.cumsum(-1)
    lo, hi = (a, b) if a < b else (b, a)
    # boundaries strictly between the two picks: cdf[lo] .. cdf[hi-1]
    edges = cdf[lo:hi]
    return float((edges - u * float(cdf[-1])).abs().min())


def check_stochastic_choices(own_dev, rec, tsp, csp, uniforms, tol=2e-3):
    """own_dev [T, G] device picks (teacher-forced run), rec = oracle record of the same forced run.  Every pick must be
    the oracle's pick from the same uniform, or the uniform must sit within `tol` of the CDF boundary between the two
    (logits agree to ~1e-3 relative, so a boundary can move by about that much).  Returns the number of boundary cases."""
    n_boundary = 0
    T, G = own_dev.shape
    for f in range(T):
        for g in range(G):
            want, got = int(rec["own_codes"][f][g]), int(own_dev[f, g])
            if want == got:
                continue
            if g == 0:
                s, sp = rec["talker_scores"][f], tsp       # processed with the forced history (penalty, masks, top-k)
            else:
                s = O.process_logits(rec["cp_logits"][f][g - 1], csp, (), g - 1)
                sp = csp
            assert torch.isfinite(s[got]), f"frame {f} group {g}: device sampled a filtered id {got}"
            d = draw_boundary_distance(s, sp, uniforms(f, g), want, got)
            assert d <= tol, f"frame {f} group {g}: device {got} vs oracle {want}, uniform is {d:.3e} away from the boundary"
            n_boundary += 1
    return n_boundary
