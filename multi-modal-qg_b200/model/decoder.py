"""B200 drop-in for the reference's model/decoder.py (AttnDecoder, Decoder).

Same constructor / forward signatures, attributes and state_dict keys as
/root/reference/model/decoder.py; arithmetic in libmmqg.so via mmqg.functional.  CUDA only.
"""
import torch
from torch.nn import Module, LSTM, Linear, Dropout
from torch.nn.init import xavier_uniform_, orthogonal_, normal_

from mmqg import functional as MF


def _init_lstm(lstm):
    for param in lstm.parameters():
        if len(param.shape) >= 2:
            orthogonal_(param.data)
        else:
            normal_(param.data)


class Decoder(Module):
    """Non-attention decoder (reference decoder.py:7-47; only used by the dead non_attn_train.py)."""

    def __init__(self, num_layers, dropout, hidden_dim, n_vocab, word_emb_dim, av_emb_dim, emb_layer):
        super().__init__()
        self.num_layers = num_layers
        self.dropout = dropout
        self.hidden_dim = hidden_dim
        self.n_vocab = n_vocab
        self.word_emb_dim = word_emb_dim
        self.av_emb_dim = av_emb_dim
        self.word_embeddings = emb_layer
        self.lstm = LSTM(self.word_emb_dim + self.av_emb_dim, self.hidden_dim, self.num_layers, dropout=self.dropout)
        self.out_layer = Linear(self.hidden_dim, self.n_vocab)
        self.initialise_weights()

    def forward(self, text, av_enc_out, hidden):
        MF.require_cuda(text, av_enc_out)
        T = text.shape[1]
        word_emb = MF.Embedding.apply(self.word_embeddings.weight, text.reshape(-1)).view(T, -1)
        word_av_emb = torch.cat((word_emb, av_enc_out.repeat(T, 1)), dim=1)
        lstm_out, hidden = MF.lstm_stack(word_av_emb.view(T, 1, -1), hidden, self.lstm, self.training)
        logits = MF.Linear.apply(lstm_out.view(T, -1), self.out_layer.weight, self.out_layer.bias).view(T, 1, -1)
        return logits, hidden

    def init_state(self, batch_sz):
        dev = self.out_layer.weight.device
        return (torch.zeros(self.num_layers, batch_sz, self.hidden_dim, device=dev),
                torch.zeros(self.num_layers, batch_sz, self.hidden_dim, device=dev))

    def initialise_weights(self):
        _init_lstm(self.lstm)
        xavier_uniform_(self.out_layer.weight)
        normal_(self.out_layer.bias)


class AttnDecoder(Module):
    """One decoder step (reference decoder.py:49-125): embedding, three location-attention heads
    over the text / audio / video memories, 3-layer LSTM step, vocabulary projection."""

    def __init__(self, num_layers, dropout_p, hidden_dim, n_vocab, word_emb_dim, video_emb_dim, audio_emb_dim, emb_layer,
                 text_max_length, av_max_length, device):
        super(AttnDecoder, self).__init__()
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.n_vocab = n_vocab
        self.dropout_p = dropout_p
        self.text_max_length = text_max_length
        self.av_max_length = av_max_length
        self.video_emb_dim = video_emb_dim
        self.audio_emb_dim = audio_emb_dim
        self.word_emb_dim = word_emb_dim
        self.emb_layer = emb_layer
        self.device = device
        self.text_attn = Linear(self.word_emb_dim + self.hidden_dim, self.text_max_length)
        self.vid_attn = Linear(self.word_emb_dim + self.hidden_dim, self.av_max_length)
        self.audio_attn = Linear(self.word_emb_dim + self.hidden_dim, self.av_max_length)
        self.dropout = Dropout(self.dropout_p)          # defined but never applied, as in the reference (Q15)
        self.lstm = LSTM(self.word_emb_dim + self.hidden_dim + self.audio_emb_dim + self.video_emb_dim, self.hidden_dim,
                         self.num_layers, dropout=self.dropout_p)
        self.out_layer = Linear(self.hidden_dim, self.n_vocab)
        self.initialise_weights()

    def forward(self, word, enc_frames, enc_seq_len, audio_emb, video_emb, hidden, encoder_outputs):
        MF.require_cuda(word, hidden[0], encoder_outputs)
        B = hidden[0].shape[1]
        TM, AM = self.text_max_length, self.av_max_length
        embedded = MF.Embedding.apply(self.emb_layer.weight, word.reshape(-1))             # (B,E)   decoder.py:75
        query = torch.cat((embedded, hidden[0][-1]), 1)                                    # (B,E+H) decoder.py:78
        # the three score Linears as one product over the concatenated weights [text|audio|video]
        w_cat = MF.cat_cached((id(self), "w"), (self.text_attn.weight, self.audio_attn.weight, self.vid_attn.weight))
        b_cat = MF.cat_cached((id(self), "b"), (self.text_attn.bias, self.audio_attn.bias, self.vid_attn.bias))
        scores = MF.Linear.apply(query, w_cat, b_cat)
        # no length mask: decoder.py:79,85,93 are no-ops in the reference;
        # enc_seq_len / enc_frames only bound the context sums (rows beyond them are zero padding)
        M_txt = encoder_outputs.view(B, TM, -1)
        M_aud = audio_emb.reshape(B, AM, -1).to(torch.float32)
        M_vid = video_emb.view(B, AM, -1)
        attn, ctx = MF.Attention.apply(scores, M_txt, M_aud, M_vid, int(enc_seq_len), int(enc_frames))
        output = torch.cat((embedded, ctx), 1).unsqueeze(0)                                 # decoder.py:99-101
        output, hidden = MF.lstm_stack(output, hidden, self.lstm, self.training)            # decoder.py:104
        logits = MF.Linear.apply(output[0], self.out_layer.weight, self.out_layer.bias)     # decoder.py:106
        text_attn_weights = attn[:, :TM]
        audio_attn_weights = attn[:, TM:TM + AM]
        vid_attn_weights = attn[:, TM + AM:TM + 2 * AM]
        return logits, hidden, text_attn_weights, audio_attn_weights, vid_attn_weights

    def initialise_weights(self):
        _init_lstm(self.lstm)
        xavier_uniform_(self.out_layer.weight)
        normal_(self.out_layer.bias)
        xavier_uniform_(self.text_attn.weight)
        normal_(self.text_attn.bias)
        xavier_uniform_(self.audio_attn.weight)
        normal_(self.audio_attn.bias)
        xavier_uniform_(self.vid_attn.weight)
        normal_(self.vid_attn.bias)
