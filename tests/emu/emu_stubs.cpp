// (every kernel family is built into the emulated library now; nothing to stub)
