// HDF5 records of Molecule / BasisSet (src/external/hdf5_read.cxx:47-156, hdf5_write.cxx in the reference,
// which go through HighFive + libhdf5) and dense FP64 datasets, through a self-contained reader / writer of
// the HDF5 subset those files use: superblock v0, v1 object headers, symbol-table groups, contiguous (or
// compact) datasets, compound types of fixed-size members.  No libhdf5 in this image.
#pragma once
#include "types.hpp"
#include <string>
#include <vector>

namespace GauXC {

void read_hdf5_record(Molecule& mol, const std::string& fname, const std::string& dset);
void read_hdf5_record(BasisSet& basis, const std::string& fname, const std::string& dset);
void write_hdf5_record(const Molecule& mol, const std::string& fname, const std::string& dset);
void write_hdf5_record(const BasisSet& basis, const std::string& fname, const std::string& dset);

// dense FP64 dataset (e.g. /DENSITY, /VXC, /EXC of the reference's fixtures); dims in file order
void read_hdf5_dataset(const std::string& fname, const std::string& dset, std::vector<double>& data,
                       std::vector<size_t>& dims);
void write_hdf5_dataset(const std::string& fname, const std::string& dset, const double* data,
                        const std::vector<size_t>& dims);

}  // namespace GauXC
