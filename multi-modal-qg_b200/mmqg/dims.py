"""Shape bookkeeping for the hot path (symbols table).

Names follow the reference's config.py (`context_max_lenth`, `av_max_length`, ...)
and its constructor arguments (reference model/encoder.py:81, model/decoder.py:50).
"""
from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class Dims:
    B: int            # samples in the batch (reference: 1, train.py:233)
    T_t: int          # context tokens per sample            (train.py:164 context_len)
    T_v: int          # salient frames == salient audio clips (train.py:155 n_frames)
    T_q: int          # question tokens incl. <end>           (train.py:171 target_len)
    V: int            # vocabulary size                       (train.py:236 n_vocab)
    E: int = 300      # GloVe dim                             (config.py:55)
    H: int = 512      # text/decoder LSTM hidden              (config.py:80,85)
    L: int = 3        # text/decoder LSTM layers              (config.py:81,86)
    H_a: int = 128    # VGGish embedding dim                  (config.py:66)
    H_v: int = 512    # video LSTM hidden                     (config.py:76)
    F_v: int = 1000   # frame feature dim                     (config.py:77 flatten_dim)
    TM: int = 283     # text attention slots                  (config.py:70 context_max_lenth)
    AM: int = 101     # audio/video attention slots           (config.py:71 av_max_length)

    def __post_init__(self):
        assert self.T_t <= self.TM and self.T_v <= self.AM
        assert self.B >= 1 and self.T_t >= 1 and self.T_v >= 1 and self.T_q >= 1

    @property
    def S(self):      # all attention slots: text | video | audio (our packing order)
        return self.TM + 2 * self.AM

    @property
    def Q(self):      # attention query width  [emb ; h_top]  (decoder.py:64-66)
        return self.E + self.H

    @property
    def C(self):      # context width [c_txt ; c_aud ; c_vid] (decoder.py:99)
        return self.H + self.H_a + self.H_v

    @property
    def X0(self):     # decoder LSTM layer-0 input width      (decoder.py:69)
        return self.E + self.C

    def asdict(self):
        return asdict(self)


# The five BASELINE.json configs.
def config(n: int, B: int = None) -> Dims:
    if n in (1, 2, 3, 5):
        b = {1: 16, 2: 256, 3: 256, 5: 1024}[n]
        return Dims(B=B or b, T_t=100, T_v=10, T_q=30 if n == 5 else 20, V=10000,
                    F_v=2048, TM=283, AM=101)
    if n == 4:
        return Dims(B=B or 256, T_t=400, T_v=64, T_q=20, V=50000, F_v=2048, TM=400, AM=101)
    raise ValueError(n)


def param_shapes(d: Dims) -> dict:
    """Flat parameter dictionary: name -> shape.  Names are the reference modules'
    state_dict keys under the prefixes emb. / text. / video. / dec."""
    G = 4 * d.H
    s = {"emb.weight": (d.V, d.E)}
    for l in range(d.L):
        i = d.E if l == 0 else d.H
        s[f"text.lstm.weight_ih_l{l}"] = (G, i)
        s[f"text.lstm.weight_hh_l{l}"] = (G, d.H)
        s[f"text.lstm.bias_ih_l{l}"] = (G,)
        s[f"text.lstm.bias_hh_l{l}"] = (G,)
    s["video.lstm.weight_ih_l0"] = (4 * d.H_v, d.F_v)
    s["video.lstm.weight_hh_l0"] = (4 * d.H_v, d.H_v)
    s["video.lstm.bias_ih_l0"] = (4 * d.H_v,)
    s["video.lstm.bias_hh_l0"] = (4 * d.H_v,)
    s["dec.text_attn.weight"] = (d.TM, d.Q)
    s["dec.text_attn.bias"] = (d.TM,)
    s["dec.vid_attn.weight"] = (d.AM, d.Q)
    s["dec.vid_attn.bias"] = (d.AM,)
    s["dec.audio_attn.weight"] = (d.AM, d.Q)
    s["dec.audio_attn.bias"] = (d.AM,)
    for l in range(d.L):
        i = d.X0 if l == 0 else d.H
        s[f"dec.lstm.weight_ih_l{l}"] = (G, i)
        s[f"dec.lstm.weight_hh_l{l}"] = (G, d.H)
        s[f"dec.lstm.bias_ih_l{l}"] = (G,)
        s[f"dec.lstm.bias_hh_l{l}"] = (G,)
    s["dec.out_layer.weight"] = (d.V, d.H)
    s["dec.out_layer.bias"] = (d.V,)
    return s
