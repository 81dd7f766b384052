"""`URMPDataset` as imported by scripts/resynthesise_dataset.py:9.  The reference's data/urmp.py
defines only URMPDataModule, so that script cannot start as shipped; the
dataset it means has GeneralDataset's three-argument constructor (data/general.py:10)."""
from .general import GeneralDataset


class URMPDataset(GeneralDataset):
    pass
