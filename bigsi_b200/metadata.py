"""SampleMetadata: colour <-> sample-name maps (mirror of bigsi/graph/metadata.py).

Host-side only: the reference stores these as `metadata:<k>` keys in
the KV store; here they live in a plain dict owned by the resident index store.
"""
DELETION_SPECIAL_SAMPLE_NAME = "D3L3T3D"  # bigsi/graph/metadata.py:1


class SampleMetadata:
    def __init__(self, storage):
        # `storage` is any mutable mapping; keys mirror the reference's "metadata:<key>" schema
        self.storage = storage

    colour_count_key = "colour_count"

    @property
    def num_samples(self):
        return self.storage.get(self._key(self.colour_count_key), 0)

    def add_sample(self, sample_name):
        self._validate_sample_name(sample_name)
        colour = self.num_samples
        self.storage[self._key(sample_name)] = colour
        self.storage[self._key(colour)] = sample_name
        self.storage[self._key(self.colour_count_key)] = colour + 1
        return colour + 1  # the reference returns the NEW count (metadata.py:22-27)

    def add_samples(self, sample_names):
        for s in sample_names:
            self.add_sample(s)

    def delete_sample(self, sample_name):
        colour = self.sample_to_colour(sample_name)
        self.storage[self._key(colour)] = DELETION_SPECIAL_SAMPLE_NAME
        self.storage[self._key(sample_name)] = -1

    def sample_name_exists(self, sample_name):
        return isinstance(self.storage.get(self._key(sample_name)), int)

    def sample_to_colour(self, sample_name):
        c = self.storage.get(self._key(sample_name))
        if not isinstance(c, int) or c < 0:
            return None
        return c

    def colour_to_sample(self, colour):
        return self.storage[self._key(colour)]

    def samples_to_colours(self, sample_names):
        return {s: self.sample_to_colour(s) for s in sample_names if self.sample_to_colour(s) is not None}

    def colours_to_samples(self, colours):
        return {c: self.colour_to_sample(c) for c in colours if self.colour_to_sample(c)}

    def sample_names(self):
        return [self.colour_to_sample(c) for c in range(self.num_samples)]

    @staticmethod
    def _key(key):
        # sample names and colours share one namespace in the reference ("metadata:<x>"); keep
        # ints and strings apart so a sample called "0" cannot shadow colour 0's entry
        return ("metadata", key)

    def _validate_sample_name(self, sample_name):
        if sample_name == DELETION_SPECIAL_SAMPLE_NAME:
            raise ValueError("You can't call a sample %s" % DELETION_SPECIAL_SAMPLE_NAME)
        if self.sample_name_exists(sample_name):
            raise ValueError("You can't insert two samples with the same name")
