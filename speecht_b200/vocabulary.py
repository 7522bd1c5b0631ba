"""Character vocabulary of the acoustic model: the same id assignment as the reference (speecht/vocabulary.py:16-21):
ids 0-25 = a-z, 26 = apostrophe, 27 = space; 28 symbols, and the CTC blank is class 28 (speech_model.py:301).

Table-driven: one alphabet string defines both directions.
"""
ALPHABET = "abcdefghijklmnopqrstuvwxyz' "

SIZE = len(ALPHABET)
APOSTROPHE = ALPHABET.index("'")
SPACE_ID = ALPHABET.index(' ')
A_ASCII_CODE = ord('a')

_TO_ID = {symbol: index for index, symbol in enumerate(ALPHABET)}


def letter_to_id(letter):
  """Vocabulary id of one character (a-z, apostrophe, space); raises KeyError for anything else."""
  return _TO_ID[letter]


def id_to_letter(identifier):
  """Character of one vocabulary id."""
  return ALPHABET[int(identifier)]


def sentence_to_ids(sentence):
  """Lower-cases `sentence` and encodes it character by character."""
  return [_TO_ID[symbol] for symbol in sentence.lower()]


def ids_to_sentence(identifiers):
  """Inverse of sentence_to_ids for any iterable of ids (ints or numpy integers)."""
  return ''.join(ALPHABET[int(identifier)] for identifier in identifiers)
