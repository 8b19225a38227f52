// YAML_Doc.hpp -- root of the report tree (interface of the reference's YAML_Doc.hpp:108-133).
#ifndef HPCCG_B200_YAML_DOC_HPP
#define HPCCG_B200_YAML_DOC_HPP

#include <string>

#include "YAML_Element.hpp"

class YAML_Doc : public YAML_Element {
 public:
  YAML_Doc(const std::string &miniApp_Name, const std::string &miniApp_Version,
           const std::string &destination_Directory = "", const std::string &destination_FileName = "");
  ~YAML_Doc();
  // Returns the document text and also writes it to
  // <dir>/<file or "name-version_">YYYY_MM_DD__HH_MM_SS.yaml (YAML_Doc.cpp:32-72).
  std::string generateYAML();

 protected:
  std::string miniAppName;
  std::string miniAppVersion;
  std::string destinationDirectory;
  std::string destinationFileName;
};

#endif
