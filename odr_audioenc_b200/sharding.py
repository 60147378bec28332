"""Host-side work partitioning for multi-GPU runs: no collective on the data path.

* one long stream -> contiguous frame ranges, one per rank; each range needs a PCM halo before it and one look-ahead
  frame after it (the DAB ScF-CRC of frame n holds frame n+1's scalefactors: toolame.c:527-542);
* an ensemble of services -> whole services per rank, longest-processing-time-first by a bitrate-weighted cost.
"""
from collections import namedtuple

FrameRange = namedtuple("FrameRange", "f0 f1 history_samples has_next")

SAMPLES_PER_FRAME = 1152
HALO_SAMPLES = 480  # polyphase history (subband.c:211-215); psy-1's 192 samples lie inside it


def time_shards(n_frames, world_size, halo=SAMPLES_PER_FRAME):
    """Split frames [0, n_frames) into world_size contiguous ranges (sizes differ by at most one frame).
    `history_samples` = PCM samples before the range the rank must be given (0 at the stream start)."""
    if halo < HALO_SAMPLES:
        raise ValueError("halo must cover the %d-sample polyphase history" % HALO_SAMPLES)
    out, base, extra = [], n_frames // world_size, n_frames % world_size
    f0 = 0
    for r in range(world_size):
        f1 = f0 + base + (1 if r < extra else 0)
        out.append(FrameRange(f0, f1, min(f0 * SAMPLES_PER_FRAME, halo), f1 < n_frames and f1 > f0))
        f0 = f1
    return out


def pcm_slice(rng, nch=None):
    """(first_sample, end_sample) of the interleaved-PCM rows a rank needs for its range."""
    first = rng.f0 * SAMPLES_PER_FRAME - rng.history_samples
    end = (rng.f1 + (1 if rng.has_next else 0)) * SAMPLES_PER_FRAME
    return first, end


def service_cost(sample_rate, nch, bitrate_kbps, n_frames):
    """Relative encode cost of a service: analysis work scales with channels and frames, allocation/packing with
    the bit budget."""
    return n_frames * (nch * 1.0 + bitrate_kbps / 384.0)


def service_shards(services, world_size):
    """services: list of (sample_rate, nch, bitrate_kbps, n_frames).  Returns per-rank lists of service indices
    (longest processing time first; ties keep input order so every rank computes the same plan)."""
    order = sorted(range(len(services)), key=lambda i: (-service_cost(*services[i]), i))
    load = [0.0] * world_size
    plan = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        plan[r].append(i)
        load[r] += service_cost(*services[i])
    return plan
