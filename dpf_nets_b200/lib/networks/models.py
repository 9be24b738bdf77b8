"""VAE glue around the fused point decoder, with the reference's class names, config keys,
`util_mode`s, output-dict keys and state_dict layout (lib/networks/models.py:13-464).

forward(g_input, p_input[, images], n_sampled_points=None) -> dict of lists:
  g_posterior_{mus,logvars,samples}, g_prior_{samples,mus,logvars}, p_prior_{samples,mus,logvars}
with index 0 of the p_prior_* lists = base prior and the decoder's per-layer outputs after it
(training: samples = decoder list + [p_input], as in the reference)."""
import torch
import torch.nn as nn

from .decoders import GlobalRNVPDecoder, LocalCondRNVPDecoder, prepend
from .encoders import FeatureEncoder, PointNetCloudEncoder
from .resnet import resnet18


def _reparameterize(mu, logvar):
    return torch.randn_like(mu) * torch.exp(0.5 * logvar) + mu


class _DPFBase(nn.Module):
    """Shared pieces of the two model variants."""

    def _build_common(self, kwargs, with_base_var):
        k = kwargs.get
        self.mode = k('util_mode')
        self.deterministic = k('deterministic')
        self.pc_enc_init_n_channels = k('pc_enc_init_n_channels')
        self.pc_enc_init_n_features = k('pc_enc_init_n_features')
        self.pc_enc_n_features = k('pc_enc_n_features')
        self.g_latent_space_size = k('g_latent_space_size')
        self.g_prior_n_flows = k('g_prior_n_flows')
        self.g_prior_n_features = k('g_prior_n_features')
        self.g_posterior_n_layers = k('g_posterior_n_layers')
        self.p_latent_space_size = k('p_latent_space_size')
        self.p_prior_n_layers = k('p_prior_n_layers')
        self.p_decoder_n_flows = k('p_decoder_n_flows')
        self.p_decoder_n_features = k('p_decoder_n_features')
        self.p_decoder_base_type = k('p_decoder_base_type')
        if with_base_var:
            self.p_decoder_base_var = k('p_decoder_base_var')
        G, P = self.g_latent_space_size, self.p_latent_space_size
        self.pc_encoder = PointNetCloudEncoder(self.pc_enc_init_n_channels, self.pc_enc_init_n_features,
                                               self.pc_enc_n_features)
        self.g_prior = GlobalRNVPDecoder(self.g_prior_n_flows, self.g_prior_n_features, G, weight_std=0.01)
        self.g_posterior = FeatureEncoder(self.g_posterior_n_layers, self.pc_enc_n_features[-1], G,
                                          deterministic=False, mu_weight_std=0.0033, mu_bias=0.0,
                                          logvar_weight_std=0.033, logvar_bias=0.0)
        if self.p_decoder_base_type == 'free':
            self.p_prior = FeatureEncoder(self.p_prior_n_layers, G, P, deterministic=False, mu_weight_std=0.001,
                                          mu_bias=0.0, logvar_weight_std=0.01, logvar_bias=0.0)
        elif self.p_decoder_base_type == 'freevar':
            self.register_buffer('p_prior_mus', torch.zeros((1, P, 1)))
            self.p_prior = FeatureEncoder(self.p_prior_n_layers, G, P, deterministic=True, mu_weight_std=0.01, mu_bias=0.0)
        elif self.p_decoder_base_type == 'fixed':
            self.register_buffer('p_prior_mus', torch.zeros((1, P, 1)))
            self.register_buffer('p_prior_logvar', float(self.p_decoder_base_var) * torch.ones((1, P, 1)))
        self.pc_decoder = LocalCondRNVPDecoder(self.p_decoder_n_flows, self.p_decoder_n_features, G, weight_std=0.01)

    def reparameterize(self, mu, logvar):
        return _reparameterize(mu, logvar)

    def _base_prior(self, g, B, n_points):
        """Base point prior (mu0, logvar0) expanded to (B, P, n_points) as stride-0 views."""
        P = self.p_latent_space_size
        if self.p_decoder_base_type == 'free':
            mu, lv = self.p_prior(g)
            mu, lv = mu.unsqueeze(2), lv.unsqueeze(2)
        elif self.p_decoder_base_type == 'freevar':
            mu, lv = self.p_prior_mus, self.p_prior(g).unsqueeze(2)
        else:
            mu, lv = self.p_prior_mus, self.p_prior_logvar
        return mu.expand(B, P, n_points), lv.expand(B, P, n_points)

    def _posterior(self, g_input, sample):
        feats = self.pc_encoder.global_features(g_input)
        mus, logvars = self.g_posterior(feats)
        return mus, logvars, (_reparameterize(mus, logvars) if sample else mus)

    def _decode_direct(self, out, g, B, n_points):
        mu0, lv0 = self._base_prior(g, B, n_points)
        z0 = _reparameterize(mu0, lv0)
        ps, mus, lvs = self.pc_decoder(z0, g, mode='direct')
        out['p_prior_samples'] = prepend(z0, ps)
        out['p_prior_mus'] = prepend(mu0, mus)
        out['p_prior_logvars'] = prepend(lv0, lvs)

    def _encode_paths(self, out, g_input, p_input, g0_mu, g0_lv, n_sampled, training):
        B = g_input.shape[0]
        mus, logvars, g = self._posterior(g_input, sample=training)
        out['g_posterior_mus'], out['g_posterior_logvars'], out['g_posterior_samples'] = mus, logvars, g
        gs, gmus, glvs = self.g_prior(g, mode='inverse')
        out['g_prior_samples'] = gs + [g]
        out['g_prior_mus'] = [g0_mu] + gmus
        out['g_prior_logvars'] = [g0_lv] + glvs
        if training:
            mu0, lv0 = self._base_prior(g, B, p_input.shape[2])
            ps, pmus, plvs = self.pc_decoder(p_input, g, mode='inverse')
            out['p_prior_samples'] = ps + [p_input] if not hasattr(ps, 'stacked') else _append(ps, p_input)
            out['p_prior_mus'] = prepend(mu0, pmus)
            out['p_prior_logvars'] = prepend(lv0, plvs)
        else:
            self._decode_direct(out, g, B, n_sampled)


def _append(flow_list, last):
    out = list(flow_list)
    out.append(last)
    return out


class Local_Cond_RNVP_MC_Global_RNVP_VAE(_DPFBase):
    def __init__(self, **kwargs):
        super().__init__()
        G = kwargs.get('g_latent_space_size')
        self.g0_prior_mus = nn.Parameter(torch.empty(1, G))
        self.g0_prior_logvars = nn.Parameter(torch.empty(1, G))
        with torch.no_grad():
            self.g0_prior_mus.normal_(mean=0.0, std=0.033)
            self.g0_prior_logvars.normal_(mean=0.0, std=0.33)
        self._build_common(kwargs, with_base_var=True)

    def encode(self, g_input):
        feats = self.pc_encoder.global_features(g_input)
        return {'g_posterior_mus': self.g_posterior(feats)[0]}

    def decode(self, g_sample, n_sampled_points=2048):
        out = {}
        self._decode_direct(out, g_sample, g_sample.shape[0], n_sampled_points)
        return out

    def forward(self, g_input, p_input, n_sampled_points=None):
        from ...ops._counters import deferred
        with deferred():      # the ~35 BatchNorm step counters of the fused paths: one multi-tensor add
            return self._forward(g_input, p_input, n_sampled_points)

    def _forward(self, g_input, p_input, n_sampled_points=None):
        n = p_input.shape[2] if n_sampled_points is None else n_sampled_points
        B, G = g_input.shape[0], self.g_latent_space_size
        out = {}
        g0_mu, g0_lv = self.g0_prior_mus.expand(B, G), self.g0_prior_logvars.expand(B, G)
        if self.mode in ('training', 'evaluating'):
            self._encode_paths(out, g_input, p_input, g0_mu, g0_lv, n, training=self.mode == 'training')
        elif self.mode == 'generating':
            g0 = _reparameterize(g0_mu, g0_lv)
            gs, gmus, glvs = self.g_prior(g0, mode='direct')
            out['g_prior_samples'] = [g0] + gs
            out['g_prior_mus'] = [g0_mu] + gmus
            out['g_prior_logvars'] = [g0_lv] + glvs
            self._decode_direct(out, gs[-1], p_input.shape[0], n)
        return out


class Local_Cond_RNVP_MC_Global_RNVP_VAE_IC(_DPFBase):
    """Single-view reconstruction variant: image-conditioned latent prior (ResNet-18 features ->
    g0_prior head), `predicting` mode decodes from the image alone."""

    def __init__(self, **kwargs):
        super().__init__()
        G = kwargs.get('g_latent_space_size')
        self.g_prior_n_layers = kwargs.get('g_prior_n_layers')
        self.img_encoder = resnet18(num_classes=G)
        self.g0_prior = FeatureEncoder(self.g_prior_n_layers, G, G, deterministic=False, mu_weight_std=0.0033,
                                       mu_bias=0.0, logvar_weight_std=0.033, logvar_bias=0.0)
        self._build_common(kwargs, with_base_var=kwargs.get('p_decoder_base_type') == 'fixed')

    def forward(self, g_input, p_input, images, n_sampled_points=None):
        from ...ops._counters import deferred
        with deferred():
            return self._forward(g_input, p_input, images, n_sampled_points)

    def _forward(self, g_input, p_input, images, n_sampled_points=None):
        n = p_input.shape[2] if n_sampled_points is None else n_sampled_points
        out = {}
        g0_mu, g0_lv = self.g0_prior(self.img_encoder(images))
        if self.mode in ('training', 'evaluating'):
            self._encode_paths(out, g_input, p_input, g0_mu, g0_lv, n, training=self.mode == 'training')
        elif self.mode == 'predicting':
            gs, gmus, glvs = self.g_prior(g0_mu, mode='direct')
            out['g_prior_samples'] = [g0_mu] + gs
            out['g_prior_mus'] = [g0_mu] + gmus
            out['g_prior_logvars'] = [g0_lv] + glvs
            self._decode_direct(out, gs[-1], p_input.shape[0], n)
        return out
