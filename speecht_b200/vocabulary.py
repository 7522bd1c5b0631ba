"""Character vocabulary of the acoustic model (mirror of reference speecht/vocabulary.py:16-81).

ids 0-25 = a-z, 26 = apostrophe, 27 = space; SIZE = 28 and the CTC blank is class 28 (speech_model.py:301)."""
APOSTROPHE = 26
SPACE_ID = 27
A_ASCII_CODE = ord('a')
SIZE = 28


def letter_to_id(letter):
  if letter == ' ':
    return SPACE_ID
  if letter == '\'':
    return APOSTROPHE
  return ord(letter) - A_ASCII_CODE


def id_to_letter(identifier):
  if identifier == SPACE_ID:
    return ' '
  if identifier == APOSTROPHE:
    return '\''
  return chr(identifier + A_ASCII_CODE)


def sentence_to_ids(sentence):
  return [letter_to_id(letter) for letter in sentence.lower()]


def ids_to_sentence(identifiers):
  return ''.join(id_to_letter(int(identifier)) for identifier in identifiers)
