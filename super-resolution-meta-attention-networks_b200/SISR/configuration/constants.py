"""Path constants (reference: Code/SISR/configuration/constants.py:3-6)."""
import os

# the directory that plays the role of the reference's `Code/` folder
base_directory = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir, os.pardir))
results_directory = os.path.join(os.path.dirname(base_directory), 'Results')
data_directory = os.path.join(os.path.dirname(base_directory), 'Data')
