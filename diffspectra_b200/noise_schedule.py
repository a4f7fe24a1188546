"""VP noise schedule with the reference's interface (diffusion/noise_schedule.py:6-122).

``NoiseScheduleVP(schedule, continuous_beta_0, continuous_beta_1)``, ``.T``, ``.marginal_prob(t)`` etc. keep the
reference's names, argument meaning and op order; the sampler (sampling.py in this package) calls
``marginal_prob`` 2 x steps times ONCE at setup to build the coefficient table the CUDA loop indexes, so a
reference ``NoiseScheduleVP`` object can be passed in unchanged as well (duck-typed).
Only the continuous schedules ('cosine' — the QM9S config, configs/diffspectra_qm9s.py:40 — and 'linear') are
provided; the discrete/interpolated variants are outside the sampling hot path.
"""
import math

import torch


class NoiseScheduleVP:
    def __init__(self, schedule='cosine', betas=None, alphas_cumprod=None, continuous_beta_0=0.1,
                 continuous_beta_1=20., dtype=torch.float32):
        if schedule not in ('linear', 'cosine'):
            raise ValueError("Unsupported noise schedule {}. The schedule needs to be 'linear' or 'cosine'".format(schedule))
        self.schedule = schedule
        self.total_N = 1000
        self.beta_0 = continuous_beta_0
        self.beta_1 = continuous_beta_1
        self.cosine_s = 0.008
        self.cosine_beta_max = 999.
        self.cosine_t_max = math.atan(self.cosine_beta_max * (1. + self.cosine_s) / math.pi) * 2. * (
            1. + self.cosine_s) / math.pi - self.cosine_s
        self.cosine_log_alpha_0 = math.log(math.cos(self.cosine_s / (1. + self.cosine_s) * math.pi / 2.))
        # T = 1 has numerical issues for the cosine schedule; the reference ends at 0.9946 (noise_schedule.py:48-51)
        self.T = 0.9946 if schedule == 'cosine' else 1.

    def marginal_log_mean_coeff(self, t):
        """log(alpha_t) for a continuous-time label t in [0, T]."""
        if self.schedule == 'linear':
            return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        log_alpha_t = torch.log(torch.cos((t + self.cosine_s) / (1. + self.cosine_s) * math.pi / 2.))
        return log_alpha_t - self.cosine_log_alpha_0

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1. - torch.exp(2. * self.marginal_log_mean_coeff(t)))

    def marginal_prob(self, t):
        log_mean_coeff = self.marginal_log_mean_coeff(t)
        return torch.exp(log_mean_coeff), torch.sqrt(1. - torch.exp(2. * log_mean_coeff))

    def marginal_lambda(self, t):
        log_mean_coeff = self.marginal_log_mean_coeff(t)
        log_std = 0.5 * torch.log(1. - torch.exp(2. * log_mean_coeff))
        return log_mean_coeff - log_std

    def get_noiseLevel(self, t):
        alpha_t = self.marginal_alpha(t)
        sigma_t = self.marginal_std(t)
        return torch.log(alpha_t ** 2 / sigma_t ** 2)


def ancestral_coefficients(noise_scheduler, time_steps):
    """[steps,4] fp32 rows (c_x, c_pred, sigma, noise_level) computed with exactly the op order of
    sampling.py:571-584,605-606 on ``time_steps``' device/dtype (so fp32 rounding — e.g. alpha_s != 1 at s = 0 —
    matches the reference; SURVEY.md §7 'schedule coefficients')."""
    t_array = time_steps
    s_array = torch.cat([time_steps[1:], torch.zeros(1, device=time_steps.device, dtype=time_steps.dtype)])
    rows = []
    for i in range(len(t_array)):
        t, s = t_array[i], s_array[i]
        alpha_t, sigma_t = noise_scheduler.marginal_prob(t)
        alpha_s, sigma_s = noise_scheduler.marginal_prob(s)
        alpha_t_given_s = alpha_t / alpha_s
        sigma2_t_given_s = sigma_t ** 2 - alpha_t_given_s ** 2 * sigma_s ** 2
        sigma_t_given_s = torch.sqrt(sigma2_t_given_s)
        sigma = sigma_t_given_s * sigma_s / sigma_t
        c_x = alpha_t_given_s * sigma_s ** 2 / sigma_t ** 2
        c_pred = alpha_s * sigma2_t_given_s / sigma_t ** 2
        noise_level = torch.log(alpha_t ** 2 / sigma_t ** 2)
        rows.append(torch.stack([c_x, c_pred, sigma, noise_level]))
    return torch.stack(rows).to(torch.float32).contiguous()
