"""Train_base: the loss / optimiser library of the reference's Training/train_base.py (class Train_base
:14-154; only `_loss_GAN` and the Adam factory are ever used by Train_goodGAN.py) on the sm_100a path.

The reference composes each loss from ~15 TF ops and differentiates it with tf.gradients; here each of
d_loss / g_loss / c_loss is ONE fused kernel that also emits the closed-form dlogits (csrc/loss.cu), so
`_entropy`, `_balance_entropy`, `_softmax_cross_entropy_loss_w_logits` and
`_sigmoid_cross_entopy_w_logits` (train_base.py:43-57, 75-84) are folded into those kernels rather than
exposed as separate ops.
"""
import torch

from . import _lib, ops
from .core import ctx


class AdamOptimizer:
    """tf.train.AdamOptimizer(learning_rate, beta1) (train_base.py:91-97) over one network's flat
    parameter buffer.  lr and the beta-power accumulators live on the device (graph-replay safe)."""

    def __init__(self, lr, beta1, beta2=0.999, epsilon=1e-8, name='Adam_optimizer'):
        self.beta1, self.beta2, self.eps, self.name = beta1, beta2, epsilon, name
        self._lr = float(lr)
        self.state = None

    def _state(self):
        if self.state is None:
            self.state = torch.tensor([self._lr, self.beta1, self.beta2], dtype=torch.float32, device=ctx.device)
        return self.state

    def set_lr(self, lr):
        if float(lr) != self._lr or self.state is None:
            self._lr = float(lr)
            _lib.call('tgan_fill_f32', self._state().data_ptr(), self._lr, 1, ops._st())

    def apply_flat(self, fb, grad_scale=1.0, ema=None, ema_decay=0.9999):
        st = self._state()
        _lib.call('tgan_adam', fb['theta'].data_ptr(), fb['m'].data_ptr(), fb['v'].data_ptr(), fb['grad'].data_ptr(),
                  fb['n'], st.data_ptr(), self.beta1, self.beta2, self.eps, grad_scale,
                  None if ema is None else ema.data_ptr(), ema_decay, ops._st())
        _lib.call('tgan_adam_advance', st.data_ptr(), self.beta1, self.beta2, ops._st())


class Train_base(object):
    def __init__(self):
        pass

    def _Adam_optimizer(self, lr, beta1, name='Adam_optimizer'):
        return AdamOptimizer(lr, beta1, name=name)

    def _loss_GAN(self, D, C, Y, Lambda):
        """train_base.py:113-154 -> (d_loss, g_loss, c_loss) as ops.Loss objects.  `Lambda` is the device
        fp32 [2] tensor {lambda_1, lambda_2}."""
        D_real, D_real_logits, D_fake, D_fake_logits, D_unl, D_unl_logits = D
        if self.config.DATA_NAME == "cifar10":
            C_real_logits, C_unl_logits, C_unl_d_logits, C_fake_logits, C_unl_logits_rep = C
        else:
            C_real_logits, C_unl_logits, C_unl_d_logits, C_fake_logits = C
            C_unl_logits_rep = None
        y_g, y_l_c = Y
        d_loss = ops.loss_d(D_real_logits, D_fake_logits, D_unl_logits)
        g_loss = ops.loss_g(D_fake_logits)
        c_loss = ops.loss_c(C_real_logits, y_l_c, C_unl_logits, C_unl_logits_rep, D_unl_logits, C_fake_logits, y_g,
                            Lambda)
        return d_loss, g_loss, c_loss
