"""The SVB acoustic model (BASELINE cfg 5's middle step): drop-in for the reference's ``MleSVBVAE``
(``modules/voice_conversion/svb_vae.py:13-56,178-199,251-312``) on the inference / evaluation path the SVB task takes
(``forward(..., infer=False, concurrent_ways=['a2a', 'p2p', 'a2p'])``, ``tasks/singing/svb_vae_task.py:120-150,303-305``):

    conditions   pitch_embed -> pitch_encoder (ConvStacks) ; vc_asr (PPG extractor) -> upsample_layer ; spk_embed_proj
    a2a / p2p    encoded_embed_proj -> GlobalFVAE (posterior encoder + mel decoder + KL)
    a2p          z_mapping_function (GlobalLatentMap) -> decoder with the professional pitch / aligned amateur content -> mel_out

Same constructor (``dict_size`` + hparams) and ``state_dict`` names as the reference.  Every Conv1d / Linear runs as a native kernel
(``svb_conv_nct_forward``; the WN stacks and the PPG extractor through their own entry points, see ``modules/fastspeech/fs2_vae.py`` and
``vc_modules.py``); embedding lookup, GroupNorm / BatchNorm(eval) affine maps, nearest up-sampling, gathers and concatenations are torch
ops on the same device tensors.  No training path (the acoustic model's DDP step, BASELINE cfg 4, is not part of this package)."""
import torch
import torch.nn.functional as F
from torch import nn

from neuralsvb_b200.modules.fastspeech.fs2_vae import GlobalFVAE
from neuralsvb_b200.modules.voice_conversion.vc_modules import VCASR, _bn_affine, _conv, _Holder


def _linear_btc(x, lin):
    """nn.Linear on [B, T, C] as a native 1x1 convolution on [B, C, T]."""
    return _conv(x.transpose(1, 2).contiguous(), lin.weight, lin.bias).transpose(1, 2)


class _ConvStacks(nn.Module):
    """Parameter names of ``ConvStacks`` / ``ConvBlock(norm='gn')`` (modules/commons/common_layers.py:672-707,739-773)."""

    def __init__(self, H, n_layers=3, K=5):
        super().__init__()
        self.K = K
        self.in_proj = nn.Linear(H, H)
        self.conv = nn.ModuleList()
        for _ in range(n_layers):
            blk, cn = _Holder(), _Holder()
            cn.conv = nn.Conv1d(H, H, K, padding=K // 2)
            blk.conv, blk.norm = cn, nn.GroupNorm(H // 16, H)
            self.conv.append(blk)
        self.out_proj = nn.Linear(H, H)

    def forward(self, x):
        x = _conv(x.transpose(1, 2).contiguous(), self.in_proj.weight, self.in_proj.bias)
        for blk in self.conv:
            h = _conv(x, blk.conv.conv.weight, blk.conv.conv.bias, K=self.K, pad=self.K // 2)
            x = x + torch.relu(F.group_norm(h, blk.norm.num_groups, blk.norm.weight, blk.norm.bias, blk.norm.eps))
        return _conv(x.contiguous(), self.out_proj.weight, self.out_proj.bias).transpose(1, 2)


class _GlobalLatentMap(nn.Module):
    """``GlobalLatentMap`` (modules/voice_conversion/vae_models.py:149-172): 1x1 convs on the [B, latent, 1] latent, eval-mode BatchNorm."""

    def __init__(self, latent):
        super().__init__()
        self.convs = nn.Sequential(nn.Conv1d(latent, latent, 1), nn.BatchNorm1d(latent), nn.ReLU(), nn.Conv1d(latent, latent, 1),
                                   nn.BatchNorm1d(latent), nn.ReLU(), nn.Conv1d(latent, latent, 1))
        self.spk_proj = nn.Sequential(nn.Conv1d(256, latent, 1), nn.ReLU(), nn.Conv1d(latent, latent, 1))

    def forward(self, x, spk_emb):
        s = spk_emb[:, :, :x.shape[-1]].contiguous()
        s = _conv(_conv(s, self.spk_proj[0].weight, self.spk_proj[0].bias, relu=True), self.spk_proj[2].weight, self.spk_proj[2].bias)
        x = (x + s).contiguous()
        c = self.convs
        x = torch.relu(_bn_affine(c[1], _conv(x, c[0].weight, c[0].bias))).contiguous()
        x = torch.relu(_bn_affine(c[4], _conv(x, c[3].weight, c[3].bias))).contiguous()
        return _conv(x, c[6].weight, c[6].bias)


class MleSVBVAE(nn.Module):
    def __init__(self, dict_size=None, hp=None, precision='bf16x3'):
        super().__init__()
        if hp is None:
            from neuralsvb_b200.utils.hparams import hparams as hp
        H = self.hidden_size = hp['hidden_size']
        self.c_content = self.c_out = hp['audio_num_mel_bins']
        self.vae_model = GlobalFVAE(self.c_content, hp['fvae_enc_dec_hidden'], hp['latent_size'], hp['fvae_kernel_size'],
                                    hp['fvae_enc_n_layers'], hp['fvae_dec_n_layers'], H, [4], False, precision=precision)
        self.pitch_embed = nn.Embedding(300, H, 0)
        self.pitch_encoder = _ConvStacks(H, 3)
        self.vc_asr = VCASR(dict_size, self.c_content, hidden_size=H, asr_enc_layers=hp['asr_enc_layers'], mel_strides=hp['mel_strides'],
                            asr_last_norm=hp.get('asr_last_norm', True))
        self.mel_strides = list(hp['mel_strides'])
        self.upsample_layer = nn.Sequential(*([nn.Sequential(nn.Upsample(scale_factor=s, mode='nearest'), nn.Conv1d(H, H, s * 2 + 1, padding=s),
                                                            nn.ReLU(), nn.BatchNorm1d(H)) for s in self.mel_strides if s > 1] +
                                              [nn.Conv1d(H, H, 5, padding=2)]))
        self.spk_embed_proj = nn.Linear(256, H)
        self.encoded_embed_proj = nn.Linear(3 * H, H)
        self.z_mapping_function = _GlobalLatentMap(hp['latent_size'])

    def load_state_dict(self, state_dict, strict=True):
        sd = {k: v for k, v in state_dict.items() if not k.startswith(('vc_asr.asr_decoder.', 'vc_asr.token_embed.'))}
        return super().load_state_dict(sd, strict=strict)

    def prepare_condition(self, mels_content=None, pitch=None, spk_ids=None):
        """svb_vae.py:57-84."""
        T = pitch.shape[1]
        h_pitch = self.pitch_encoder(F.embedding(pitch, self.pitch_embed.weight))
        h = self.vc_asr(mels_content)['h_content'].transpose(1, 2)
        for m in self.upsample_layer:
            if isinstance(m, nn.Sequential):                        # Upsample(nearest) -> Conv1d -> ReLU -> BatchNorm1d (eval)
                s = int(m[0].scale_factor)
                h = _bn_affine(m[3], _conv(h.repeat_interleave(s, dim=2).contiguous(), m[1].weight, m[1].bias, K=2 * s + 1, pad=s, relu=True))
            else:
                h = _conv(h.contiguous(), m.weight, m.bias, K=5, pad=2)
        h_content = h.transpose(1, 2)[:, :mels_content.shape[1]]
        h_style = _linear_btc(spk_ids[:, None, :].float(), self.spk_embed_proj).repeat(1, T, 1)
        return {'h_pitch': h_pitch, 'h_content': h_content, 'h_style': h_style, 'tgt_nonpadding': (pitch > 0).float()[:, :, None]}

    def _cond_sum(self, parts):
        return _linear_btc(torch.cat(parts, -1), self.encoded_embed_proj).transpose(1, 2).contiguous()

    def normal_vae(self, tgt_mel, pitch_cond, content_cond, timbre_cond, padding_cond, infer, eps=None):
        """svb_vae.py:155-165 (infer=False: the branch the SVB task uses for validation and test as well)."""
        if infer:
            raise NotImplementedError('normal_vae(infer=True) samples a [B, latent, T / 4] prior that the reference GlobalFVAEDecoder cannot '
                                      'consume (vae_models.py:120-122); the SVB task always calls it with infer=False')
        g = self._cond_sum([pitch_cond, content_cond, timbre_cond])
        x_recon, kl, z_p, m_q, logs_q, x_mask_sqz, z_q = self.vae_model(tgt_mel.transpose(1, 2).contiguous(), padding_cond.transpose(1, 2), g, eps=eps)
        return {'mel_out': x_recon.transpose(1, 2), 'kl': kl, 'z_p': z_p, 'm_q': m_q, 'logs_q': logs_q, 'x_mask_sqz': x_mask_sqz, 'z_q': z_q}

    def forward(self, amateur_mel=None, prof_mel=None, amateur_pitch=None, prof_pitch=None, amateur_spk_id=None, prof_spk_id=None,
                a2p_alignment=None, p2a_alignment=None, infer=False, disable_map=False, eps=None, **kwargs):
        """svb_vae.py:258-312.  ``eps`` [B, latent, 1] replaces the posterior noise (parity tests); default torch.randn_like."""
        if torch.is_grad_enabled() and self.training:
            raise RuntimeError('neuralsvb_b200 MleSVBVAE is inference only: call .eval() and run under torch.no_grad()')
        ways = kwargs['concurrent_ways']
        ret = {}
        ac = self.prepare_condition(amateur_mel, amateur_pitch, spk_ids=amateur_spk_id)
        pc = self.prepare_condition(prof_mel, prof_pitch, spk_ids=prof_spk_id)
        if 'a2a' in ways:
            a2a = ret['a2a'] = self.normal_vae(amateur_mel, ac['h_pitch'], ac['h_content'], ac['h_style'], ac['tgt_nonpadding'], infer, eps)
        if 'p2p' in ways:
            p2p = ret['p2p'] = self.normal_vae(prof_mel, pc['h_pitch'], pc['h_content'], pc['h_style'], pc['tgt_nonpadding'], infer, eps)
        if 'a2p' in ways:
            z = a2a['z_q']
            mapped = z if disable_map else self.z_mapping_function(z, ac['h_style'].transpose(1, 2))
            prof = torch.distributions.Normal(p2p['m_q'], p2p['logs_q'].exp(), validate_args=False)      # no host sync
            out = {'mle': -prof.log_prob(mapped).sum() / mapped.shape[0] / mapped.shape[1]}
            align = a2p_alignment[:, :, None].repeat(1, 1, self.hidden_size)
            g = self._cond_sum([pc['h_pitch'], torch.gather(ac['h_content'], 1, align),
                                ac['h_style'][:, :1, :].repeat(1, pc['h_pitch'].shape[1], 1)])
            out['mel_out'] = self.vae_model.decoder(mapped.contiguous(), pc['tgt_nonpadding'].transpose(1, 2), g).transpose(1, 2)
            out['logs_amateur_zq'], out['logs_prof_zq'] = a2a['z_q'], p2p['z_q']
            ret['a2p'] = out
        return ret
