"""Config: the class-attribute configuration object of the reference (config.py:9-92) -- same
attribute names -- plus the three per-dataset TempConfig presets that the reference defines inline in
Training/Train_goodGAN.py (:484-530 svhn, :560-607 cifar10, :641-685 mnist).  Only attributes that the
training hot path reads are kept; logging / checkpoint / summary switches are out of scope."""


class Config(object):
    NAME = None
    DATA_NAME = None
    DATA_DIR = None
    NUM_LABEL = None
    BATCH_SIZE = None
    BATCH_SIZE_L_D = None
    SAMPLE_SIZE = 64
    IMAGE_HEIGHT = None
    IMAGE_WIDTH = None
    CHANNEL = None
    Z_DIM = None
    NUM_CLASSES = None
    MINIBATCH_DIS = False
    BATCH_NORM_DECAY = 0.9
    BATCH_NORM_EPSILON = 1e-5
    LEARNING_RATE = 3e-4
    CLA_LEARNINIG_RATE = 3e-4
    BETA1 = 0.5
    FAKE_G_LAMBDA = 0.0
    PRE_TRAIN = False
    EPOCHS = None
    TRAIN_SIZE = None

    def __init__(self):
        """Set values of computed attributes."""
        self.IMAGE_DIM = [self.IMAGE_HEIGHT, self.IMAGE_WIDTH, self.CHANNEL]

    def config_str(self):
        s = "\nConfigurations:\n"
        for a in dir(self):
            if not a.startswith("__") and not callable(getattr(self, a)):
                s += "{:30} {}\n".format(a, getattr(self, a))
        return s

    def display(self):
        print(self.config_str())


class Cifar10Config(Config):       # Train_goodGAN.py:560-607
    NAME = "Good_GAN"
    DATA_NAME = "cifar10"
    NUM_LABEL = 4000
    BATCH_SIZE_G = 100
    BATCH_SIZE_L_C = 50
    BATCH_SIZE_U_C = 50
    BATCH_SIZE_L_D = 20
    BATCH_SIZE_U_D = 80
    BATCH_SIZE = BATCH_SIZE_G
    IMAGE_HEIGHT, IMAGE_WIDTH, CHANNEL = 32, 32, 3
    FAKE_G_LAMBDA = 0.3
    Z_DIM = 100
    NUM_CLASSES = 10
    LEARNING_RATE = 3e-4
    CLA_LEARNINIG_RATE = 3e-3
    EPOCHS = 1000
    TRAIN_SIZE = 60000 - NUM_LABEL
    ZCA = None                     # (mean[3072], mat[3072,3072]) when DATA_DIR holds no cifar10_zca_*.npy


class SvhnConfig(Config):          # Train_goodGAN.py:484-530
    NAME = "Good_GAN"
    DATA_NAME = "svhn"
    NUM_LABEL = 500
    BATCH_SIZE = 100
    BATCH_SIZE_G = BATCH_SIZE
    BATCH_SIZE_L_C = 50
    BATCH_SIZE_U_C = 50
    BATCH_SIZE_L_D = 20
    BATCH_SIZE_U_D = 80
    IMAGE_HEIGHT, IMAGE_WIDTH, CHANNEL = 32, 32, 3
    FAKE_G_LAMBDA = 0.03
    CLA_LEARNINIG_RATE = 3e-4
    Z_DIM = 100
    NUM_CLASSES = 10
    LEARNING_RATE = 3e-4
    EPOCHS = 1000
    TRAIN_SIZE = 73257 - NUM_LABEL


class MnistConfig(Config):         # Train_goodGAN.py:641-685
    NAME = "Good_GAN"
    DATA_NAME = "mnist"
    NUM_LABEL = 100
    BATCH_SIZE_G = 100
    BATCH_SIZE_L_C = 100
    BATCH_SIZE_U_C = 100
    BATCH_SIZE_L_D = 20
    BATCH_SIZE_U_D = 80
    BATCH_SIZE = BATCH_SIZE_G
    IMAGE_HEIGHT, IMAGE_WIDTH, CHANNEL = 28, 28, 1
    FAKE_G_LAMBDA = 0.1
    Z_DIM = 100
    NUM_CLASSES = 10
    LEARNING_RATE = 1e-3
    CLA_LEARNINIG_RATE = 3e-4
    EPOCHS = 1000
    TRAIN_SIZE = 60000 - NUM_LABEL


def make_config(data_name, scale=1, **over):
    """Preset for `data_name`; `scale` divides every batch size (small parity cases)."""
    cls = {'cifar10': Cifar10Config, 'svhn': SvhnConfig, 'mnist': MnistConfig}.get(data_name)
    if cls is None:
        raise ValueError("The specified dataset is not yet implemented!")
    cfg = cls()
    for k in ('BATCH_SIZE', 'BATCH_SIZE_G', 'BATCH_SIZE_L_C', 'BATCH_SIZE_U_C', 'BATCH_SIZE_L_D', 'BATCH_SIZE_U_D'):
        setattr(cfg, k, getattr(cfg, k) // scale)
    for k, v in over.items():
        setattr(cfg, k, v)
    # _main_training_svhn / _main_training_cifar10 (Train_goodGAN.py:537-538, 614-615): with fewer than 1000 labels
    # the first 30 epochs train the classifier alone; the mnist main has the rule commented out (:692-693)
    if data_name in ('svhn', 'cifar10') and cfg.NUM_LABEL < 1000 and 'PRE_TRAIN' not in over:
        cfg.PRE_TRAIN = True
    return cfg
