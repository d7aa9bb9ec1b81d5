"""reference: Code/sr_tools/helper_functions.py (create_dir_if_empty :161-164, read_metadata :129-138)."""
import csv
import os


def create_dir_if_empty(*directories):
    for d in directories:
        if not os.path.exists(d):
            os.mkdir(d)


def read_metadata(file):
    """one-line csv with a header -> dict"""
    with open(file) as fh:
        return next(csv.DictReader(fh))
