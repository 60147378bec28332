"""Host-side work partitioning for multi-GPU runs: no collective on the data path.

* one long stream -> contiguous frame ranges, one per rank; each range needs a PCM halo before it and one look-ahead
  frame after it (the DAB ScF-CRC of frame n holds frame n+1's scalefactors: toolame.c:527-542);
* an ensemble of services -> whole services per rank (`service_shards`, longest-processing-time-first), or
  services cut in time as well (`ensemble_shards`) when whole services cannot balance: 18 equal services on
  8 GPUs leave one rank with 3 of them (6.0x at best), cut in time every rank gets 18/8 of a service.
"""
from collections import namedtuple

FrameRange = namedtuple("FrameRange", "f0 f1 history_samples has_next")
Piece = namedtuple("Piece", "service f0 f1 history_samples has_next")

SAMPLES_PER_FRAME = 1152
HALO_SAMPLES = 480        # polyphase history (subband.c:211-215); psy-1's 192 samples lie inside it
HALO_SAMPLES_PSY2 = 1632  # psy model 2: FFT window of the block two before the frame's first (psycho_2.c:80-92)


def halo_for(psy_model=1, halo_samples=None):
    """PCM history a mid-stream range must be given, in whole frames' worth of samples: what tlb_info.halo_samples
    reports (480, or 1632 with psy model 2), rounded up to a multiple of 1152 so that shards stay frame aligned."""
    need = halo_samples if halo_samples is not None else (HALO_SAMPLES_PSY2 if psy_model == 2 else HALO_SAMPLES)
    if need < HALO_SAMPLES:
        raise ValueError("halo must cover the %d-sample polyphase history" % HALO_SAMPLES)
    return -(-need // SAMPLES_PER_FRAME) * SAMPLES_PER_FRAME


def _range(f0, f1, n_frames, halo):
    return FrameRange(f0, f1, min(f0 * SAMPLES_PER_FRAME, halo), f1 < n_frames and f1 > f0)


def time_shards(n_frames, world_size, halo=None, psy_model=1, halo_samples=None):
    """Split frames [0, n_frames) into world_size contiguous ranges (sizes differ by at most one frame).
    `history_samples` = PCM samples before the range the rank must be given (0 at the stream start).
    Pass the encoder's psy model (or tlb_info.halo_samples) so that the history covers what the model looks back
    on; `halo` overrides the rounded value and must itself cover it."""
    need = halo_for(psy_model, halo_samples)
    if halo is None:
        halo = need
    elif halo < (halo_samples if halo_samples is not None else (HALO_SAMPLES_PSY2 if psy_model == 2 else HALO_SAMPLES)):
        raise ValueError("halo %d is shorter than the %d samples psy model %d looks back on" % (halo, need, psy_model))
    out, base, extra = [], n_frames // world_size, n_frames % world_size
    f0 = 0
    for r in range(world_size):
        f1 = f0 + base + (1 if r < extra else 0)
        out.append(_range(f0, f1, n_frames, halo))
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


def ensemble_shards(services, world_size, psy_model=1, halo_samples=None, min_piece_frames=64):
    """Balance an ensemble by service AND by time: the services are laid end to end on a cost axis (grouped by
    configuration so that a rank sees few distinct encoders) and cut into world_size stretches of equal cost; a
    service that straddles a cut is split at a frame boundary, each piece carrying its PCM halo and look-ahead frame
    like a time shard.  Returns per-rank lists of Piece(service, f0, f1, history_samples, has_next); every frame of
    every service appears exactly once; at most world_size - 1 services are cut."""
    halo = halo_for(psy_model, halo_samples)
    order = sorted(range(len(services)), key=lambda i: (services[i][:3], i))
    per_frame = [service_cost(*services[i][:3], 1) for i in range(len(services))]
    total = sum(per_frame[i] * services[i][3] for i in order)
    plan = [[] for _ in range(world_size)]
    if total <= 0:
        return plan
    cuts = [total * k / world_size for k in range(1, world_size)]  # cost positions where rank k begins
    c0, rank = 0.0, 0
    for i in order:
        n = services[i][3]
        c1 = c0 + n * per_frame[i]
        # frame positions of the cuts that fall inside this service; a cut that would leave a sliver (a piece costs
        # its halo and look-ahead frames on top) moves to the service's nearer end
        marks = []
        for k in range(rank, world_size - 1):
            if cuts[k] >= c1:
                break
            f = int(round((cuts[k] - c0) / per_frame[i]))
            f = 0 if f < min_piece_frames else (n if n - f < min_piece_frames else f)
            marks.append(f)
        f0 = 0
        for f in marks:
            f = max(f, f0)
            if f > f0:
                rng = _range(f0, f, n, halo)
                plan[rank].append(Piece(i, rng.f0, rng.f1, rng.history_samples, rng.has_next))
            f0 = f
            rank += 1
        if n > f0:
            rng = _range(f0, n, n, halo)
            plan[rank].append(Piece(i, rng.f0, rng.f1, rng.history_samples, rng.has_next))
        c0 = c1
    return plan


def plan_cost(plan, services):
    """Per-rank cost of an ensemble_shards plan (for reporting and for the balance tests)."""
    return [sum(service_cost(*services[p.service][:3], p.f1 - p.f0) for p in pieces) for pieces in plan]
