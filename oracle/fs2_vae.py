"""ORACLE (test infrastructure only; see oracle/__init__.py): torch-CPU restatement of the WN gated dilated-conv
stack, modules/fastspeech/fs2_vae.py:11-94 of the reference.  Pinned by tests/golden/wn.npz, which
oracle/gen_golden.py writes from the UNMODIFIED reference class."""
import torch
import torch.nn.functional as F


def fold_weight_norm(sd):
    """WN.remove_weight_norm (fs2_vae.py:96-103): weight = g * v / ||v||, norm over every dim except 0."""
    out = {}
    for k, v in sd.items():
        if k.endswith('.weight_g'):
            continue
        if k.endswith('.weight_v'):
            out[k[:-2]] = torch._weight_norm(v, sd[k[:-1] + 'g'], 0)
        else:
            out[k] = v
    return out


def wn_forward(w, hidden, kernel_size, dilation_rate, n_layers, x, x_mask=None, g=None):
    """WN.forward (fs2_vae.py:62-94) on folded weights ``w``; x [B, H, T], x_mask [B, 1, T] or None, g [B, gin, T] or None."""
    mask = 1 if x_mask is None else x_mask
    output = torch.zeros_like(x)
    if g is not None:
        g = F.conv1d(g, w['cond_layer.weight'], w['cond_layer.bias'])                        # :73-74
    for i in range(n_layers):
        d = dilation_rate ** i
        x_in = F.conv1d(x, w[f'in_layers.{i}.weight'], w[f'in_layers.{i}.bias'], dilation=d,
                        padding=int((kernel_size * d - d) / 2))                                # :43-46,77
        if g is not None:
            x_in = x_in + g[:, i * 2 * hidden:(i + 1) * 2 * hidden]                           # :80-82
        acts = torch.tanh(x_in[:, :hidden]) * torch.sigmoid(x_in[:, hidden:])                # :11-17
        rs = F.conv1d(acts, w[f'res_skip_layers.{i}.weight'], w[f'res_skip_layers.{i}.bias'])  # :88
        if i < n_layers - 1:
            x = (x + rs[:, :hidden]) * mask                                                   # :90
            output = output + rs[:, hidden:]                                                  # :91
        else:
            output = output + rs                                                              # :93
    return output * mask                                                                      # :94


def fvae_decoder_forward(w, hidden, kernel_size, n_layers, stride, x, x_mask, g, global_latent=False):
    """FVAEDecoder.forward (fs2_vae.py:147-152); with ``global_latent`` GlobalFVAEDecoder.forward (vae_models.py:120-128).
    ``w``: folded state_dict (pre_net.0.*, wn.*, out_proj.*)."""
    if global_latent:
        x = x.repeat(1, 1, g.shape[-1] // stride)
    x = F.conv_transpose1d(x, w['pre_net.0.weight'], w['pre_net.0.bias'], stride=stride)
    x = x * x_mask
    wn_w = {k[3:]: v for k, v in w.items() if k.startswith('wn.')}
    x = wn_forward(wn_w, hidden, kernel_size, 1, n_layers, x, x_mask if torch.is_tensor(x_mask) else None, g) * x_mask
    return F.conv1d(x, w['out_proj.weight'], w['out_proj.bias'])


def global_fvae_encoder_forward(w, hidden, latent, kernel_size, n_layers, stride, x, x_mask, g, eps):
    """GlobalFVAEEncoder.forward (vae_models.py:96-106) in eval mode on a folded state_dict; ``eps`` replaces randn_like(m)."""
    x = F.conv1d(x, w['pre_net.0.weight'], w['pre_net.0.bias'], stride=stride, padding=stride // 2)
    x_mask = x_mask[:, :, ::stride][:, :, :x.shape[-1]]
    x = x * x_mask
    wn_w = {k[3:]: v for k, v in w.items() if k.startswith('wn.')}
    x = wn_forward(wn_w, hidden, kernel_size, 1, n_layers, x, x_mask, g) * x_mask
    x = F.conv1d(x, w['out_proj.weight'], w['out_proj.bias'])
    for i in (0, 3):
        x = torch.relu(F.conv1d(x, w[f'poolings.{i}.weight'], w[f'poolings.{i}.bias'], stride=2))
        j = i + 2
        x = F.batch_norm(x, w[f'poolings.{j}.running_mean'], w[f'poolings.{j}.running_var'], w[f'poolings.{j}.weight'], w[f'poolings.{j}.bias'],
                         training=False, eps=1e-5)
    x = F.conv1d(x, w['poolings.6.weight'], w['poolings.6.bias'], stride=2)
    x = torch.mean(x, dim=-1, keepdim=True)
    m, logs = torch.split(x, latent, dim=1)
    return m + eps * torch.exp(logs), m, logs, x_mask


def global_fvae_forward(w, in_out, hidden, latent, kernel_size, enc_layers, dec_layers, stride, x, x_mask, g, eps):
    """GlobalFVAE.forward(..., infer=False) (vae_models.py:11-44 via :130-146) in eval mode, ``eps`` in place of randn_like."""
    g_sqz = F.conv1d(g, w['g_pre_net.0.weight'], w['g_pre_net.0.bias'], stride=stride, padding=stride // 2)
    enc = {k[8:]: v for k, v in w.items() if k.startswith('encoder.')}
    dec = {k[8:]: v for k, v in w.items() if k.startswith('decoder.')}
    z_q, m_q, logs_q, xm = global_fvae_encoder_forward(enc, hidden, latent, kernel_size, enc_layers, stride, x, x_mask, g_sqz, eps)
    x_recon = fvae_decoder_forward(dec, hidden, kernel_size, dec_layers, stride, z_q, x_mask, g, True)
    kl = torch.distributions.kl_divergence(torch.distributions.Normal(m_q, logs_q.exp()), torch.distributions.Normal(0, 1))
    return x_recon, (kl * xm).sum() / xm.sum() / z_q.shape[1], m_q, logs_q
