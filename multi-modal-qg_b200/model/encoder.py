"""B200 drop-in for the reference's model/encoder.py.

Same classes, constructor arguments (order and names), forward signatures, attributes and
state_dict keys as /root/reference/model/encoder.py; the LSTM arithmetic
runs in libmmqg.so.  torch.nn.LSTM objects are kept purely as parameter containers so that
`state_dict()` / `load_state_dict()` / optimizers see exactly the reference's tensors; they are
never called.  CUDA only: CPU tensors raise (no fallback).
"""
import torch
from torch.nn import Module, LSTM, Conv2d, MaxPool2d, AdaptiveAvgPool1d, BatchNorm2d, Flatten
from torch.nn.init import orthogonal_, normal_

from mmqg import functional as MF


def _init_lstm(lstm):                       # reference encoder.py:73-78,102-107
    for param in lstm.parameters():
        if len(param.shape) >= 2:
            orthogonal_(param.data)
        else:
            normal_(param.data)


class AudioEncoder(Module):
    """VGGish feature extractor (reference encoder.py:8-19).  Upstream of the hot path and needs
    the network (torch.hub); kept only so the class exists under the same name."""

    def __init__(self):
        super().__init__()
        self.vggish = torch.hub.load('harritaylor/torchvggish', 'vggish', postprocess=False)
        self.adapt_avg_pool = AdaptiveAvgPool1d(1)

    def forward(self, audio_file):
        return self.vggish.forward(audio_file)


class VideoConvLstmEncoder(Module):
    """reference encoder.py:31-78.  Conv2d / BatchNorm2d / MaxPool2d objects are parameter and buffer containers
    (same state_dict keys); the conv stack runs on csrc/convstack.cu, the LSTM over the
    per-frame features on the recurrent kernels."""

    def __init__(self, in_channels, kernel_sz, stride, hidden_dim, video_emb_dim):
        super().__init__()
        self.in_channels = in_channels
        self.kernel_sz = kernel_sz
        self.stride = stride
        self.hidden_dim = hidden_dim
        self.video_emb_dim = video_emb_dim
        self.conv1 = Conv2d(self.in_channels, 4, self.kernel_sz, self.stride)
        self.bn1 = BatchNorm2d(4)
        self.conv2 = Conv2d(4, 6, self.kernel_sz, self.stride)
        self.bn2 = BatchNorm2d(6)
        self.maxpool1 = MaxPool2d(self.kernel_sz, self.kernel_sz)
        self.conv3 = Conv2d(6, 8, self.kernel_sz, self.stride)
        self.bn3 = BatchNorm2d(8)
        self.conv4 = Conv2d(8, 10, self.kernel_sz, self.stride)
        self.bn4 = BatchNorm2d(10)
        self.maxpool2 = MaxPool2d(self.kernel_sz, self.kernel_sz)
        self.flatten = Flatten()
        self.lstm = LSTM(self.video_emb_dim, self.hidden_dim)
        self.initialise_weights()

    def encode_features(self, feats):
        """(T, F_v) per-frame features -> (T, 1, H): the LSTM of encoder.py:69 (zero initial state)."""
        MF.require_cuda(feats)
        T = feats.shape[0]
        z = torch.zeros(1, 1, self.hidden_dim, device=feats.device, dtype=torch.float32)
        out, _ = MF.lstm_stack(feats.view(T, 1, -1), (z, z), self.lstm, self.training)
        return out

    def forward(self, video_frames):
        batch_sz = video_frames.shape[2]
        channels = video_frames.shape[1]
        height = video_frames.shape[3]
        width = video_frames.shape[4]
        # `view`, not permute: the reference reinterprets memory here
        x = video_frames.view(batch_sz, channels, height, width)
        # reference encoder.py:62-63: maxpool1(bn2(relu(conv2(bn1(relu(conv1(x))))))) and the same with conv3/4, bn3/4,
        # maxpool2 -- on the kernels of csrc/convstack.cu (train-mode batch statistics, running-stat updates included)
        MF.require_cuda(x)
        second_block = MF.conv_stack(x, self)
        cnn_out = self.flatten(second_block)
        return self.encode_features(cnn_out)

    def initialise_weights(self):
        _init_lstm(self.lstm)


class TextEncoder(Module):
    """reference encoder.py:80-111: shared embedding lookup + num_layers LSTM, one call per token
    (train.py:164-166) or per sequence; batch 1 like the reference when `text` is 0-d/1-d, batched
    (B,) tokens / (L,B,H) state as an extension."""

    def __init__(self, num_layers, dropout_p, hidden_dim, emb_dim, emb_layer, device):
        super().__init__()
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.embedding_dim = emb_dim
        self.word_embeddings = emb_layer
        self.device = device
        self.dropout_p = dropout_p
        self.lstm = LSTM(self.embedding_dim, self.hidden_dim, self.num_layers, dropout=self.dropout_p)
        self.initialise_weights()

    def forward(self, text, hidden):
        MF.require_cuda(text, hidden[0])
        B = hidden[0].shape[1]
        tokens = text.reshape(-1)
        embeds = MF.Embedding.apply(self.word_embeddings.weight, tokens)        # encoder.py:96
        T = tokens.numel() // B
        # reference: batch 1, view(L,1,E); batched extension: tokens (B,) for one step
        x = embeds.view(T, B, -1) if B == 1 else embeds.view(B, T, -1).transpose(0, 1).contiguous()
        lstm_out, hidden = MF.lstm_stack(x, hidden, self.lstm, self.training)    # encoder.py:98
        return lstm_out, hidden

    def initialise_weights(self):
        _init_lstm(self.lstm)

    def init_state(self, batch_sz):
        return (torch.zeros(self.num_layers, batch_sz, self.hidden_dim, device=self.device),
                torch.zeros(self.num_layers, batch_sz, self.hidden_dim, device=self.device))


class AudioVideoEncoder(Module):
    """reference encoder.py:113-131.  `audio_encoder_factory` lets offline callers supply a
    feature pass-through in place of the network-bound VGGish model."""
    audio_encoder_factory = AudioEncoder

    def __init__(self, av_in_channels, av_kernel_sz, av_stride, av_hidden_dim, video_emb_dim):
        super().__init__()
        self.audio_enc = type(self).audio_encoder_factory()
        self.video_enc = VideoConvLstmEncoder(av_in_channels, av_kernel_sz, av_stride, av_hidden_dim, video_emb_dim)

    def forward(self, audio_file, video_frames):
        audio_out = self.audio_enc(audio_file)
        # The reference flattens to (1, n*128) here (encoder.py:123), which its own decoder cannot
        # consume for n > 1; the layout the decoder's bmm needs is (n, 128).
        audio_emb = audio_out.view(-1, audio_out.shape[-1])
        video_emb = self.video_enc(video_frames).squeeze(1)
        return audio_emb, video_emb
