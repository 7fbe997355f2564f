"""Good_GAN: the MNIST / SVHN (/ CIFAR-10 variant) Triple-GAN builders of the reference's
Model/Good_GAN.py (class Good_GAN :10-479) on the sm_100a path.  Same method names and argument order;
the reference's svhn and cifar10 branches are layer-for-layer identical and share one code path here.
"""
from . import model_base, nn, ops
from .core import ctx, variable_scope
from .good_gan_cifar10 import _LazySigmoid, concat_batch


class Good_GAN(model_base.NN_Base):
    def __init__(self, config):
        super(Good_GAN, self).__init__(config.BATCH_NORM_DECAY, config.BATCH_NORM_EPSILON)
        self.config = config

    def _check(self):
        if self.config.DATA_NAME not in ("mnist", "svhn", "cifar10"):
            raise ValueError("The specified dataset is not yet implemented!")

    def good_generator(self, z, y, reuse=False, tag='G'):
        self._check()
        with variable_scope("good_generator", reuse=reuse):
            if self.config.DATA_NAME == "mnist":          # Good_GAN.py:19-33
                z = ops.concat_label(z, y)
                z = self._linear_fc(z, 500, 'gg_h0_lin')
                h0 = nn.softplus(z, 'gg_sp0')
                h0 = self._batch_norm_contrib(h0, 'gg_bn0', train=True)

                h1 = ops.concat_label(h0, y)
                h1 = self._linear_fc(h1, 500, 'gg_h1_lin')
                h1 = nn.softplus(h1, 'gg_sp1')
                h1 = self._batch_norm_contrib(h1, 'gg_bn1', train=True)

                h2 = ops.concat_label(h1, y)
                h2 = self._WN_dense(h2, 28 * 28, 'gg_h2_lin')
                return nn.sigmoid(h2, 'gg_sp1')
            # svhn :35-58 / cifar10 :59-83
            yb = ops.reshape(y, [y.shape[0], 1, 1, self.config.NUM_CLASSES])
            z = ops.concat_label(z, y)

            z = self._linear_fc(z, 4 * 4 * 512, 'gg_h0_lin')
            # reference order: reshape -> relu (Good_GAN.py:41-42).  relu commutes with the reshape, and the
            # fc bias is per flat feature, so bias+relu run fused on [N, 8192] and the view follows.
            h0 = nn.relu(z, 'gg_rl0')
            h0 = ops.reshape(h0, [-1, 4, 4, 512])  # [4,4]
            h0 = self._batch_norm_contrib(h0, 'gg_bn0', train=True)
            h0 = self._conv_cond_concat(h0, yb)

            h0 = self._deconv2d(h0, 256, k_w=5, k_h=5, d_w=2, d_h=2, name='gg_dconv0')  # [8, 8]
            h0 = nn.relu(h0, 'gg_rl1')
            h0 = self._batch_norm_contrib(h0, 'gg_bn1', train=True)
            h0 = self._conv_cond_concat(h0, yb)

            h1 = self._deconv2d(h0, 128, k_w=5, k_h=5, d_w=2, d_h=2, name='gg_dconv1')
            h1 = nn.relu(h1, 'gg_rl2')  # [16,16]
            h1 = self._batch_norm_contrib(h1, 'gg_bn2', train=True)
            h1 = self._conv_cond_concat(h1, yb)

            h2 = self._WN_deconv2d(h1, 3, k_w=5, k_h=5, d_w=2, d_h=2, init_scale=0.1, init=False,
                                   name='gg_wndconv0')
            return nn.tanh(h2)  # [32, 32]

    def discriminator(self, image, y, reuse=False, tag='D'):
        self._check()
        with variable_scope("discriminator", reuse=reuse):
            if self.config.DATA_NAME == "mnist":          # Good_GAN.py:93-124
                image = ops.reshape(image, [-1, 28 * 28])
                image = self._add_noise(image, stddev=0.2, tag=tag + '/noise0')
                h = ops.concat_label(image, y)
                for i in range(5):
                    h = self._WN_dense(h, [1000, 500, 250, 250, 250][i], 'd_h%d_wndense0' % i, init=False)
                    h = nn.leaky_relu(h)
                    h = self._add_noise(h, stddev=0.2, tag=tag + '/noise%d' % (i + 1))
                    h = ops.concat_label(h, y)
                h5 = self._WN_dense(h, 1, 'd_h5_wndense0', init=False)
                h5 = ops.force(h5)
                return _LazySigmoid(h5), h5
            # svhn :126-165 / cifar10 :167-206
            image = self._drop_out(image, 0.2, True, tag=tag + '/drop0')
            yb = ops.reshape(y, [image.shape[0], 1, 1, self.config.NUM_CLASSES])
            image = self._conv_cond_concat(image, yb)

            h0 = self._WN_conv2d(image, 32, k_h=3, k_w=3, d_h=1, d_w=1, init=False, name="d_h0_wnconv0")
            h0 = nn.leaky_relu(h0)
            h0 = self._conv_cond_concat(h0, yb)

            h0 = self._WN_conv2d(h0, 32, k_h=3, k_w=3, d_h=2, d_w=2, init=False, name="d_h0_wnconv1")
            h0 = nn.leaky_relu(h0)
            h0 = self._drop_out(h0, 0.2, True, tag=tag + '/drop1')  # [16, 16]

            h1 = self._conv_cond_concat(h0, yb)
            h1 = self._WN_conv2d(h1, 64, k_h=3, k_w=3, d_h=1, d_w=1, init=False, name="d_h1_wnconv0")
            h1 = nn.leaky_relu(h1)
            h1 = self._conv_cond_concat(h1, yb)

            h1 = self._WN_conv2d(h1, 64, k_h=3, k_w=3, d_h=2, d_w=2, init=False, name="d_h1_wnconv1")
            h1 = nn.leaky_relu(h1)
            h1 = self._drop_out(h1, 0.2, True, tag=tag + '/drop2')  # [8, 8]

            h2 = self._conv_cond_concat(h1, yb)
            h2 = self._WN_conv2d(h2, 128, k_h=3, k_w=3, d_h=1, d_w=1, init=False, name="d_h2_wnconv0")
            h2 = nn.leaky_relu(h2)
            h2 = self._conv_cond_concat(h2, yb)

            h2 = self._conv_cond_concat(h2, yb)       # the double concat of Good_GAN.py:151-153 -> 148 channels
            h2 = self._WN_conv2d(h2, 128, k_h=3, k_w=3, d_h=1, d_w=1, init=False, name="d_h2_wnconv1")
            h2 = nn.leaky_relu(h2)

            h3 = ops.global_pool(h2, 'mean')          # tf.reduce_mean(axis=[1, 2])
            h3 = ops.concat_label(h3, y)
            if self.config.MINIBATCH_DIS:
                raise NotImplementedError('MINIBATCH_DIS is off in every reference config (Train_goodGAN.py:511)')
            h3 = self._WN_dense(h3, 1, 'd_h3_wndense')
            h3 = ops.force(h3)
            return _LazySigmoid(h3), h3

    def classifier(self, image, train_ph, reuse=False, tag='C'):
        self._check()
        tr = bool(train_ph)

        def blk(x, conv, bn, cout):
            h = self._conv2d(x, cout, k_h=3, k_w=3, d_h=1, d_w=1, name=conv)
            h = nn.leaky_relu(h)
            return self._batch_norm_contrib(h, name=bn, train=tr)

        with variable_scope("classifier", reuse=reuse):
            if self.config.DATA_NAME == "mnist":          # Good_GAN.py:216-247
                image = ops.reshape(image, [-1, 28, 28, 1])
                image = self._add_noise(image, stddev=0.3, tag=tag + '/noise')
                h0 = blk(image, 'c_h0_conv0', 'c_h0_bn0', 32)
                h0 = ops.max_pool2(h0)
                h0 = self._drop_out(h0, 0.5, tr, tag=tag + '/drop1')
                h1 = blk(h0, 'c_h1_conv0', 'c_h1_bn0', 64)
                h1 = blk(h1, 'c_h1_conv1', 'c_h1_bn1', 64)
                h1 = ops.max_pool2(h1)
                h1 = self._drop_out(h1, 0.5, tr, tag=tag + '/drop2')
                h2 = blk(h1, 'c_h2_conv0', 'c_h2_bn0', 128)
                h2 = blk(h2, 'c_h2_conv1', 'c_h2_bn1', 128)
            else:                                          # svhn :249-299 / cifar10 :301-350
                image = self._drop_out(image, 0.2, tr, tag=tag + '/drop0')
                h0 = image
                for i in range(3):
                    h0 = blk(h0, 'c_h0_conv%d' % i, 'c_h0_bn%d' % i, 128)
                h0 = ops.max_pool2(h0)
                h0 = self._drop_out(h0, 0.5, tr, tag=tag + '/drop1')
                h1 = h0
                for i in range(3):
                    h1 = blk(h1, 'c_h1_conv%d' % i, 'c_h1_bn%d' % i, 256)
                h1 = ops.max_pool2(h1)
                h1 = self._drop_out(h1, 0.5, tr, tag=tag + '/drop2')
                h2 = blk(h1, 'c_h2_conv0', 'c_h2_bn0', 512)
                h2 = self._nin(h2, 256, name='c_h2_nin0')
                h2 = nn.leaky_relu(h2)
                h2 = self._batch_norm_contrib(h2, name='c_h2_bn1', train=tr)
                h2 = self._nin(h2, 128, name='c_h2_nin1')
                h2 = nn.leaky_relu(h2)
                h2 = self._batch_norm_contrib(h2, name='c_h2_bn2', train=tr)
            h2 = ops.global_pool(h2, 'mean')               # Global pooling
            fm = h2
            h2 = self._linear_fc(h2, self.config.NUM_CLASSES, 'c_h2_lin')
            h2 = self._batch_norm_contrib(h2, name='c_h3_bn0', train=tr)
            return h2, fm

    def good_sampler(self, z, y, reuse=True):
        return self.good_generator(z, y, reuse=reuse, tag='sampler')

    def forward_pass(self, z_g, y_g, x_l_c, y_l_c, x_l_d, y_l_d, x_u_d, x_u_c, train, tag='F'):
        """Good_GAN.forward_pass (Good_GAN.py:428-472)."""
        ex = ctx.store.has
        G = self.good_generator(z_g, y_g, reuse=ex('good_generator/gg_h0_lin/gg_h0_lin/kernel'), tag=tag + '/G')
        C_real_logits, _ = self.classifier(x_l_c, train, reuse=ex('classifier/c_h0_conv0/c_h0_conv0/kernel'),
                                           tag=tag + '/C_real')
        C_unl_logits, _ = self.classifier(x_u_c, train, reuse=True, tag=tag + '/C_unl')
        _, C_unl_onehot = ops.argmax_onehot(C_unl_logits, self.config.NUM_CLASSES)
        C_unl_d_logits, _ = self.classifier(x_u_d, train, reuse=True, tag=tag + '/C_unl_d')
        _, C_unl_d_onehot = ops.argmax_onehot(C_unl_d_logits, self.config.NUM_CLASSES)
        C_fake_logits, _ = self.classifier(G, train, reuse=True, tag=tag + '/C_fake')

        X_P = concat_batch([x_l_d, x_u_d])
        Y_P = concat_batch([y_l_d, C_unl_d_onehot])
        d_first = 'discriminator/d_h0_wndense0/V' if self.config.DATA_NAME == 'mnist' else \
            'discriminator/d_h0_wnconv0/V'
        D_real, D_real_logits = self.discriminator(X_P, Y_P, reuse=ex(d_first), tag=tag + '/D_real')
        D_fake, D_fake_logits = self.discriminator(G, y_g, reuse=True, tag=tag + '/D_fake')
        D_unl, D_unl_logits = self.discriminator(x_u_c, C_unl_onehot, reuse=True, tag=tag + '/D_unl')
        return [G, [D_real, D_real_logits, D_fake, D_fake_logits, D_unl, D_unl_logits],
                [C_real_logits, C_unl_logits, C_unl_d_logits, C_fake_logits]]

    def forward_pass_CGAN(self, z, image, y):
        G = self.good_generator(z, y, reuse=False)
        D_real, D_real_logits = self.discriminator(image, y, reuse=False)
        D_fake, D_fake_logits = self.discriminator(G, y, reuse=True)
        return G, [D_real, D_real_logits, D_fake, D_fake_logits]
