# Configuration module in the reference's format (evidence/examples/51Peg/config_51Peg_example.py):
# configdicts = [rundict, input_dict, datadict]; parameter = [init, free flag, [Prior, *shape]]
import os
from pathlib import Path

import numpy as np

here = Path(__file__).parent.absolute()

rundict = {
    'target': 'synth',
    'runid': 'b200',
    'star_params': {'star_mass': (1.0, 0.05)},
    'save_dir': os.path.join(here, 'chains'),
}

datadict = {
    inst: {'datafile': os.path.join(here, f'{inst}.rv'), 'instrument': inst,
           'kwargs': {'sep': '\t', 'skiprows': (1,)}}
    for inst in ('inst0', 'inst1')
}

planetdict1 = {'k1': [0.0, 1, ['Uniform', 0., 20.]],
               'period': [0.0, 1, ['Jeffreys', 1., 1000.]],
               'ecc': [0.1, 1, ['Beta', 0.867, 3.03]],
               'omega': [0.1, 1, ['Uniform', 0., 2 * np.pi]],
               'ma0': [0.1, 1, ['Uniform', 0., 2 * np.pi]],
               'epoch': [52500, 0]}

input_dict = {'planet1': planetdict1,
              'inst0': {'offset': [0., 1, ['Uniform', -10, 10]], 'jitter': [0.75, 1, ['Uniform', 0., 10.]]},
              'inst1': {'offset': [0., 1, ['Uniform', -10, 10]], 'jitter': [0.75, 1, ['Uniform', 0., 10.]]},
              'drift': {'lin': [0., 1, ['Uniform', -1, 1]], 'tref': [52500, 0]}}

configdicts = [rundict, input_dict, datadict]
