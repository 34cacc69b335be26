"""Drop-in for the reference's config.py (config.py:1-58): the global DEVICE and the three static
hyper-parameter classes, same names and values.  Shapes of BASELINE.json configs 1-4 come from here."""
import torch

global DEVICE
DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class RegConfig:
    save_dir = './saved_models'
    train_size = 1024
    batch_size = 128
    lr = 1e-3
    epochs = 1000
    train_samples = 5
    test_samples = 10
    num_test_points = 400
    mode = 'regression'
    mixture_prior = False
    hidden_units = 400
    noise_tolerance = .1
    mu_init = [-0.2, 0.2]
    rho_init = [-5, -4]
    prior_init = [1]
    regression_clusters = False


class RLConfig:
    data_dir = 'data/agaricus-lepiota.data'
    batch_size = 64
    num_batches = 64
    buffer_size = batch_size * num_batches
    lr = 1e-4
    training_steps = 50000
    mode = 'regression'
    hidden_units = 100
    mixture_prior = True
    mu_init = [-0.2, 0.2]
    rho_init = [-5, -4]
    prior_init = [0.5, -0, -6]


class ClassConfig:
    batch_size = 128
    lr = 1e-4
    epochs = 300
    hidden_units = 1200
    mode = 'classification'
    train_samples = 2
    test_samples = 10
    x_shape = 28 * 28
    classes = 10
    mu_init = [-0.2, 0.2]
    rho_init = [-5, -4]
    prior_init = [1.]
    mixture_prior = False
    save_dir = './saved_models'
    local_reparam = True
