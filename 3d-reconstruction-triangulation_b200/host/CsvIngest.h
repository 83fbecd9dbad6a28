// CsvIngest.h -- fast path for DetectionsContainer::readFiles (src/DetectionsContainer.cpp:19-76):
// the detection CSVs are mapped into memory and parsed with a hand-rolled integer scanner, one host
// thread per camera file, straight into the per-camera frame lists.  Same semantics as the reference's
// getline / std::stoi loop: tokens are split on ',', empty tokens skipped, every token is truncated to
// its leading integer ("0.97" -> 0), the detection is (field 5, field 6) of each record, frames whose
// number is skipped in the file become empty, a row whose token count does not fit throws
// "Invalid CSV file!".
#pragma once
#include <string>
#include <vector>

#include "DetectionsContainer.h"

// Parses `files` (camera order = order given) into out[camera][frame][detection].
void ingestDetectionFiles(const std::vector<std::string>& files, int offset, int recordSize, int startFrame, int endFrame,
                          std::vector<Cameras>& out, int n_threads = 0);
