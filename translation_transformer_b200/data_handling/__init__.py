from .tokenizer import ChemSMILESTokenizer, GenericTokenizer, split_smiles  # noqa: F401
