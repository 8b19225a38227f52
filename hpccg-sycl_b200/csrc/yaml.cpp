// yaml.cpp -- YAML_Element / YAML_Doc: the reference's report format (YAML_Element.cpp:24-120,
// YAML_Doc.cpp:17-72), written fresh.
#include <sys/stat.h>

#include <ctime>
#include <fstream>

#include "include/YAML_Doc.hpp"

YAML_Element::~YAML_Element() {
  for (YAML_Element *c : children) delete c;
}

YAML_Element *YAML_Element::attach(const std::string &k, const std::string &v) {
  value.clear();  // a key with children carries no value
  children.push_back(new YAML_Element(k, v));
  return children.back();
}

YAML_Element *YAML_Element::get(const std::string &key_arg) {
  for (YAML_Element *c : children)
    if (c->key == key_arg) return c;
  return 0;
}

std::string YAML_Element::printYAML(std::string space) {
  std::string out = space + key + ": " + value + "\n";
  const std::string deeper = space + "  ";
  for (YAML_Element *c : children) out += c->printYAML(deeper);
  return out;
}

YAML_Doc::YAML_Doc(const std::string &miniApp_Name, const std::string &miniApp_Version,
                   const std::string &destination_Directory, const std::string &destination_FileName)
    : miniAppName(miniApp_Name),
      miniAppVersion(miniApp_Version),
      destinationDirectory(destination_Directory),
      destinationFileName(destination_FileName) {}

YAML_Doc::~YAML_Doc() {}

std::string YAML_Doc::generateYAML() {
  std::string text = "Mini-Application Name: " + miniAppName + "\n" + "Mini-Application Version: " + miniAppVersion + "\n";
  for (YAML_Element *c : children) text += c->printYAML("");

  char stamp[32];
  std::time_t now = std::time(nullptr);
  std::tm tmv;
  localtime_r(&now, &tmv);
  std::strftime(stamp, sizeof stamp, "%Y_%m_%d__%H_%M_%S", &tmv);

  std::string path = (destinationFileName.empty() ? miniAppName + "-" + miniAppVersion + "_" : destinationFileName) +
                     stamp + ".yaml";
  if (!destinationDirectory.empty() && destinationDirectory != ".") {
    ::mkdir(destinationDirectory.c_str(), 0755);
    // the reference drops the timestamp when a directory is given (YAML_Doc.cpp:64)
    path = destinationDirectory + "/" + destinationFileName;
  } else {
    path = "./" + path;
  }
  std::ofstream f(path.c_str());
  f << text;
  return text;
}
